"""One launch of the spatial attention at the level-0 shape of configs[1] (28 images x 2880 tokens x 5 heads), for ncu:
ncu --set full --import-source on --clock-control none -k regex:attn_spatial -s 1 -c 1 -o gpurun_out/r2_attn python tools/attn_prof.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.ops import AttnSpatial

dev = torch.device("cuda:0")
sp = torch.cuda.current_stream().cuda_stream
n_img, S, heads = 28, 2880, 5
C = heads * 64
qkv = (torch.randn(n_img * S, 3 * C, device=dev)).to(torch.bfloat16)
out = torch.zeros(n_img * S, C, device=dev, dtype=torch.bfloat16)
op = AttnSpatial(qkv, out, n_img=n_img, heads=heads)
for _ in range(3):
    op.launch(sp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    op.launch(sp)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 5 * 1e3
print(f"attn_spatial level 0: {us:.1f} us, {op.alg_flops / us / 1e6:.0f} TFLOP/s")
