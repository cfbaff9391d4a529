"""Runs one GEMM shape a few times (for ncu captures)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from posetraj_b200.ops import Gemm  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (80640, 320, 320)
geglu = len(sys.argv) > 4 and sys.argv[4] == "geglu"
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(2 * N if geglu else N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
bias = torch.randn(w.shape[0], device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
g = Gemm(a, w, out, bias=bias, geglu=geglu)
for _ in range(4):
    g.launch(torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
