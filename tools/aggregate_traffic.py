"""Aggregates an ncu CSV holding gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum per launch of
tools/profile_step.py into (a) a single-metric launch list tools/join_launches.py can read and (b) per-kernel-class
DRAM traffic (the source of bench.py's roofline.traffic).
Usage: python tools/aggregate_traffic.py ncu.csv oplist.json out_traffic.json out_time_only.csv"""
import csv
import json
import sys
from collections import OrderedDict, defaultdict

ncu_csv, oplist_json, out_json, out_csv = sys.argv[1:5]
with open(ncu_csv) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
per_id = OrderedDict()
for r in rows:
    d = per_id.setdefault(r["ID"], {"row": r})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    name = r["Metric Name"]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    elif name == "gpu__time_duration.sum":
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}[unit]
    d[name] = v
ops = json.load(open(oplist_json))
ids = list(per_id)
assert sum(o["kernels"] for o in ops) == len(ids), (sum(o["kernels"] for o in ops), len(ids))
with open(out_csv, "w") as f:
    w = csv.DictWriter(f, fieldnames=list(rows[0].keys()), quoting=csv.QUOTE_ALL)
    w.writeheader()
    for i in ids:
        r = dict(per_id[i]["row"])
        r["Metric Name"], r["Metric Unit"], r["Metric Value"] = "gpu__time_duration.sum", "ns", str(per_id[i]["gpu__time_duration.sum"])
        w.writerow(r)
cls = defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "algorithmic_bytes": 0.0})
k = 0
for o in ops:
    for _ in range(o["kernels"]):
        d = per_id[ids[k]]
        k += 1
        c = cls[o["kind"]]
        c["launches"] += 1
        c["ms"] += d["gpu__time_duration.sum"] / 1e6
        c["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        c["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
    cls[o["kind"]]["algorithmic_bytes"] += o["bytes"]
for c in cls.values():
    c["traffic_bytes_per_launch"] = (c["dram_read_bytes"] + c["dram_write_bytes"]) / c["launches"]
json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                     "python tools/profile_step.py (one eager step, configs[1])", "classes": cls}, open(out_json, "w"), indent=1)
print(json.dumps({k: {"ms": round(v["ms"], 3), "MB/launch": round(v["traffic_bytes_per_launch"] / 1e6, 1)} for k, v in cls.items()}))
