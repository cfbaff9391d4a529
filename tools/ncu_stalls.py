"""Summarises the source page of an .ncu-rep: stall samples per SASS instruction (top N) with executed counts.
Usage: python tools/ncu_stalls.py file.ncu-rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
for k in range(len(starts) - 1):
    blk = lines[starts[k] + 1: starts[k + 1]]
    r = list(csv.reader(blk))
    hdr = r[0]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ie = hdr.index("Instructions Executed")
    rows = [x for x in r[1:] if len(x) > si]
    tot = sum(int(x[si]) for x in rows)
    print("==== kernel", k, lines[starts[k]][:120], "samples", tot, "warp-instructions", sum(int(x[ie]) for x in rows))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for x in rows:
        for c in stall_cols:
            agg[hdr[c]] = agg.get(hdr[c], 0) + int(x[c])
    print("  by reason:", {k2: v for k2, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][si]))[:topn]
    for i in sorted(top):
        x = rows[i]
        st = {hdr[c]: int(x[c]) for c in stall_cols if int(x[c]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"  {i:5d} {x[1][:72].strip():72s} smp={x[si]:>5s} exec={x[ie]:>8s} {st}")
