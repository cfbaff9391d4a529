"""Runs a few isolated GEMM launches between cudaProfilerStart/Stop (target of `ncu --profile-from-start off`).
Usage: python tools/prof_gemm.py M N K [geglu] [res]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from posetraj_b200.ops import Gemm  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
geglu = "geglu" in sys.argv
res = "res" in sys.argv
sp = torch.cuda.current_stream().cuda_stream
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn((2 * N if geglu else N), K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
bias = torch.randn(w.shape[0], device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if res else None
g = Gemm(a, w, out, bias=bias, geglu=geglu, res1=r)
for _ in range(3):
    g.launch(sp)
torch.cuda.synchronize()
torch.cuda.profiler.start()
g.launch(sp)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("block_n", g.block_n)
