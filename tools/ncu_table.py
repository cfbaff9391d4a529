"""One markdown row per kernel launch of an ncu report (raw page): time, DRAM read / write, L2->L1 bytes, tensor and MUFU
pipe activity, issue-slot utilisation, registers.  Joined with the launch-ordered op list of tools/profile_step.py when given.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv [oplist.json first_tensor_launch_index]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, body = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def col(r, name, default=""):
    i = idx.get(name)
    return r[i] if i is not None and i < len(r) else default


def num(r, name):
    try:
        return float(col(r, name).replace(",", ""))
    except ValueError:
        return float("nan")


ops = None
if len(sys.argv) > 3:
    allops = json.load(open(sys.argv[2]))
    tens = [o for o in allops if o["kind"] in ("gemm", "mlp_geglu")]
    ops = tens[int(sys.argv[3]):]
units = {h: rows[1][i] for h, i in idx.items()}
print("| # | kernel / layer | us | TFLOP/s | DRAM rd MB | DRAM wr MB | L2->L1 MB | tensor pipe % | MUFU pipe % | issue % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(body):
    if len(r) < len(hdr) // 2:
        continue
    name = col(r, "Kernel Name").split("(")[0][-60:]
    us = num(r, "gpu__time_duration.sum")
    us = us / 1e3 if units.get("gpu__time_duration.sum") in ("ns", "nsecond") else us
    def mb(metric):
        v = num(r, metric)
        u = units.get(metric, "")
        return v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
    label, tf = name, ""
    if ops is not None and n < len(ops):
        label = f"`{ops[n]['name']}`"
        tf = f"{ops[n]['flops'] / us / 1e6:.0f}" if us > 0 else ""
    print(f"| {n} | {label} | {us:.1f} | {tf} | {mb('dram__bytes_read.sum'):.0f} | {mb('dram__bytes_write.sum'):.0f} | "
          f"{mb('l1tex__m_xbar2l1tex_read_bytes.sum'):.0f} | {num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{num(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.0f} | {num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{col(r, 'launch__registers_per_thread')} |")
