"""Times the epilogue-bound GEMM shapes of configs[1] under the current PT_EPI_DEPTH / PT_EPI16 environment (CUDA events,
L2 flushed).  Usage: PT_EPI_DEPTH=2 python tools/epi_knobs.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.ops import Gemm

dev = torch.device("cuda:0")
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=9):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = []
for (M, N, K, res) in [(80640, 320, 320, 1), (80640, 320, 320, 0), (80640, 960, 320, 0), (20160, 640, 640, 1), (5040, 1280, 1280, 1),
                       (80640, 320, 1280, 1), (20160, 640, 2560, 1)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
    g = Gemm(a, w, o, bias=torch.randn(N, device=dev), res1=r)
    out.append(f"{M}x{N}x{K}{'+res' if res else ''}: {timeit(lambda: g.launch(sp)):.1f}")
print(f"EPI_DEPTH={os.environ.get('PT_EPI_DEPTH', '-')} EPI16={os.environ.get('PT_EPI16', '-')} | " + " | ".join(out))
