"""Joins an ncu launch list (gpu__time_duration.sum CSV of tools/profile_step.py) with the op list it dumped.
Usage: python tools/join_launches.py launches.csv oplist.json [out.md]"""
import csv
import json
import sys
from collections import defaultdict

launches_csv, oplist_json = sys.argv[1:3]
with open(launches_csv) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
ops = json.load(open(oplist_json))
times = [float(r["Metric Value"]) / 1e3 for r in rows]  # us
names = [r["Kernel Name"].split("(")[0] for r in rows]
assert sum(o["kernels"] for o in ops) == len(rows), (sum(o["kernels"] for o in ops), len(rows))
i = 0
per = []
for o in ops:
    t = sum(times[i:i + o["kernels"]])
    i += o["kernels"]
    per.append((o, t))
total = sum(t for _, t in per)
out = []
out.append(f"total device time of one step (ncu, serialised, cold cache): {total / 1e3:.3f} ms over {len(rows)} launches\n")
bykind = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for o, t in per:
    b = bykind[o["kind"]]
    b[0] += o["kernels"]; b[1] += t; b[2] += o["flops"]; b[3] += o["bytes"]
out.append("| kernel class | launches | ms | share | TFLOP/s (algorithmic) | GB/s (algorithmic) |\n|---|---|---|---|---|---|")
for k, b in sorted(bykind.items(), key=lambda kv: -kv[1][1]):
    tf = f"{b[2] / (b[1] * 1e-6) / 1e12:.0f}" if b[2] else "-"
    gb = f"{b[3] / (b[1] * 1e-6) / 1e9:.0f}" if b[3] else "-"
    out.append(f"| {k} | {b[0]} | {b[1] / 1e3:.3f} | {b[1] / total * 100:.1f}% | {tf} | {gb} |")
# GEMM by shape
shapes = defaultdict(lambda: [0, 0.0, 0.0])
for o, t in per:
    if o["kind"] == "gemm":
        key = (o["rows"], o["n_out"], o["k"], o["taps"], o["geglu"], o["block_n"])
        s = shapes[key]
        s[0] += 1; s[1] += t; s[2] += o["flops"]
out.append("\n| GEMM rows | N | K | taps | geglu | block_n | launches | total ms | avg us | TFLOP/s |\n|---|---|---|---|---|---|---|---|---|---|")
for key, s in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {key[0]} | {key[1]} | {key[2]} | {key[3]} | {key[4]} | {key[5]} | {s[0]} | {s[1] / 1e3:.3f} | {s[1] / s[0]:.1f} | {s[2] / (s[1] * 1e-6) / 1e12:.0f} |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text + "\n")
