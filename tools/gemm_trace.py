"""Event timeline of pt_gemm (CTA 0): MMA issuer and epilogue warp 2, per tile.  Usage: python tools/gemm_trace.py [M N K res]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200 import _lib
from posetraj_b200.ops import Gemm

dev = torch.device("cuda:0")
sp = torch.cuda.current_stream().cuda_stream
M, N, K, res = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (80640, 320, 320, 1)))
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
g = Gemm(a, w, out, bias=torch.randn(N, device=dev), res1=r)
trace = torch.zeros(8 * 64, device=dev, dtype=torch.int64)
for _ in range(2):
    g.launch(sp)
_lib.lib().pt_gemm_set_trace(trace.data_ptr())
g.launch(sp)
torch.cuda.synchronize()
_lib.lib().pt_gemm_set_trace(None)
t = trace.cpu().view(8, 64)
t0 = int(t[0, 0])
print(f"{M}x{N}x{K} res={res} block_n={g.block_n} pair={g.cta_pair}")
print("tile | MMA: wait tempty | got tempty | all k-steps issued || epilogue warp 2: wait tfull | got tfull | tile done")
for it in range(2, 8):
    print(it, [int(t[e, it]) - t0 for e in range(6)])
print("cycles per tile (epilogue):", (int(t[5, 7]) - int(t[5, 2])) / 5)
