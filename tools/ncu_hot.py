"""Hot spots of an ncu report's source page: top SASS instructions by stall samples, with the dominant stall reason.
Usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
si = idx["# Samples"]
body = [r for r in rows[2:] if len(r) > si]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[si] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {h: sum(float(r[idx[h]] or 0) for r in body) for h in stall_cols}
print("by reason:", {k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -float(body[i][si] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = max(stall_cols, key=lambda h: float(r[idx[h]] or 0))
    print(f"{i:5d} {float(r[si]):7.0f} {100 * float(r[si]) / tot:5.1f}%  {top:22s} exec={r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:90]}")
