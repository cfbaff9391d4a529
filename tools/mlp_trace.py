"""Event timeline of the fused GEGLU feed-forward (CTA 0): clock64 stamps of the MMA issuer and of gate warp 3 for the
first 64 hidden chunks.  Usage (gpurun): python tools/mlp_trace.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.ops import FusedMlp

dev = torch.device("cuda:0")
sp = torch.cuda.current_stream().cuda_stream
M, C = 80640, 320
H = 4 * C
rnd = lambda *s, scale=1.0: (torch.randn(*s, device=dev) * scale).to(torch.bfloat16)
x, w1, w2 = rnd(M, C), rnd(2 * H, C, scale=1 / math.sqrt(C)), rnd(C, H, scale=1 / math.sqrt(H))
b1, b2 = torch.randn(2 * H, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
out = torch.zeros(M, C, device=dev, dtype=torch.bfloat16)
trace = torch.zeros(2 * 8 * 64, device=dev, dtype=torch.int64)
op = FusedMlp(x, w1, b1, w2, b2, out, res1=rnd(M, C), trace=trace)
for _ in range(3):
    op.launch(sp)
torch.cuda.synchronize()
t = trace.cpu().view(2, 8, 64)
t0 = int(t[0, 0, 0])
names_m = ["wait acc1_empty", "got acc1_empty", "G1 issued", "wait p_full", "got p_full", "got w2_full", "G2 issued"]
names_e = ["wait acc1_full", "got acc1_full", "ldtm+arrive done", "math done / wait p_empty", "got p_empty", "P arrive done"]
print("chunk | MMA issuer: " + " | ".join(names_m))
for c in range(22, 30):
    print(c, [int(t[0, e, c]) - t0 for e in range(7)])
print("chunk | gate warp 3: " + " | ".join(names_e))
for c in range(22, 30):
    print(c, [int(t[1, e, c]) - t0 for e in range(6)])
per = (int(t[1, 1, 39]) - int(t[1, 1, 21])) / 18
print("cycles per chunk (chunks 21..39):", per)
