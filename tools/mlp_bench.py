"""Fused GEGLU feed-forward vs the GEGLU GEMM + output GEMM pair at the level-0 shape of configs[1] (80640 x 320, hidden
1280), CUDA events, L2 flushed between iterations.  Usage (gpurun): python tools/mlp_bench.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.ops import FusedMlp, Gemm

dev = torch.device("cuda:0")
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(torch.bfloat16)


def timeit(fn, n=10):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for M, C in ((80640, 320), (40320, 320), (20160, 256)):
    H = 4 * C
    x, w1, w2 = rnd(M, C), rnd(2 * H, C, scale=1 / math.sqrt(C)), rnd(C, H, scale=1 / math.sqrt(H))
    b1, b2 = torch.randn(2 * H, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
    res = rnd(M, C)
    out = torch.zeros(M, C, device=dev, dtype=torch.bfloat16)
    hid = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    fused = FusedMlp(x, w1, b1, w2, b2, out, res1=res)
    g1, g2 = Gemm(x, w1, hid, geglu=True, bias=b1), Gemm(hid, w2, out, bias=b2, res1=res)
    flops = fused.alg_flops
    t_f = timeit(lambda: fused.launch(sp))
    t_1 = timeit(lambda: g1.launch(sp))
    t_2 = timeit(lambda: g2.launch(sp))
    print(f"M={M} C={C}: fused {t_f:.1f} us ({flops / t_f / 1e6:.0f} TFLOP/s)   pair {t_1:.1f} + {t_2:.1f} = {t_1 + t_2:.1f} us "
          f"({flops / (t_1 + t_2) / 1e6:.0f} TFLOP/s)", flush=True)
