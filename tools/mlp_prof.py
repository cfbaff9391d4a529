"""One launch of the fused GEGLU feed-forward at the level-0 shape, for ncu (gpurun):
ncu --set full --import-source on --clock-control none -k regex:mlp_geglu -c 1 -o gpurun_out/r2_mlp python tools/mlp_prof.py"""
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
op = FusedMlp(x, w1, b1, w2, b2, out, res1=rnd(M, C))
for _ in range(3):
    op.launch(sp)
torch.cuda.synchronize()
print("done")
