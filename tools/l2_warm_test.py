"""Do small-M GEMMs speed up when their weights are already in L2?  (validates the weight-prefetch idea)"""
import math, sys, torch
sys.path.insert(0, ".")
from posetraj_b200.ops import Gemm, conv3x3_taps

sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def t(fn, warm_weights, w, iters=7):
    ts = []
    for _ in range(iters):
        flush.zero_()
        if warm_weights:
            w.float().sum()  # touch the weights: they land in L2 (126 MB)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

for (n, H, W, Cin, Cout) in [(28, 5, 9, 1280, 1280), (28, 10, 18, 1280, 1280), (28, 20, 36, 640, 640), (28, 5, 9, 2560, 1280)]:
    a = torch.randn(n * (H + 1) * (W + 1), Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device="cuda") / math.sqrt(9 * Cin)).to(torch.bfloat16)
    out = torch.empty(n * H * W, Cout, device="cuda", dtype=torch.bfloat16)
    g = Gemm(a, w, out, taps=conv3x3_taps(W), halo=(H, W))
    print(f"conv rows {a.shape[0]} {Cin}->{Cout}: cold {t(lambda: g.launch(sp), False, w):.1f} us, weights in L2 {t(lambda: g.launch(sp), True, w):.1f} us (pair={g.cta_pair} bn={g.block_n})")
for (M, N, K) in [(1260, 1280, 1280), (1260, 1280, 5120), (5040, 1280, 5120), (5040, 1280, 1280)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    g = Gemm(a, w, out)
    print(f"lin {M}x{N}x{K}: cold {t(lambda: g.launch(sp), False, w):.1f} us, weights in L2 {t(lambda: g.launch(sp), True, w):.1f} us")
