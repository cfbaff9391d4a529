"""Micro-benchmark of the tcgen05 GEMM on representative PoseTraj shapes (CUDA events, L2-flushed)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from posetraj_b200.ops import Gemm, conv3x3_taps  # noqa: E402


def bench(fn, iters=10):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    sp = torch.cuda.current_stream().cuda_stream
    rows = []
    # linear layers: (M, N, K, geglu)
    for (M, N, K, geglu) in [(80640, 320, 320, False), (80640, 960, 320, False), (80640, 1280, 320, True),
                             (80640, 320, 1280, False), (20160, 640, 640, False), (20160, 2560, 640, True),
                             (20160, 640, 2560, False), (5040, 1280, 1280, False), (5040, 5120, 1280, True),
                             (5040, 1280, 5120, False), (8192, 8192, 8192, False)]:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn((2 * N if geglu else N), K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(w.shape[0], device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        g = Gemm(a, w, out, bias=bias, geglu=geglu)
        ms = bench(lambda: g.launch(sp))
        fl = 2.0 * M * w.shape[0] * K
        t0 = bench(lambda: torch.matmul(a, w.t()))
        rows.append((f"linear M{M} N{N} K{K} geglu{int(geglu)} bn{g.block_n}", ms, fl / ms / 1e9, fl / t0 / 1e9))
        if not geglu and N <= 1280 and M > 8192:
            res = torch.randn(M, N, device="cuda").to(torch.bfloat16)
            g2 = Gemm(a, w, out, bias=bias, res1=res)
            ms = bench(lambda: g2.launch(sp))
            rows.append((f"  + residual operand", ms, fl / ms / 1e9, 0.0))
    for (n, H, W, Cin, Cout) in [(28, 40, 72, 320, 320), (28, 20, 36, 640, 640), (28, 10, 18, 1280, 1280),
                                 (28, 5, 9, 1280, 1280), (28, 10, 18, 2560, 1280)]:
        a = torch.randn(n * (H + 1) * (W + 1), Cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(Cout, 9 * Cin, device="cuda") / math.sqrt(9 * Cin)).to(torch.bfloat16)
        bias = torch.randn(Cout, device="cuda")
        out = torch.empty(n * H * W, Cout, device="cuda", dtype=torch.bfloat16)
        g = Gemm(a, w, out, taps=conv3x3_taps(W), bias=bias, halo=(H, W))
        ms = bench(lambda: g.launch(sp))
        fl = 2.0 * n * H * W * Cout * 9 * Cin
        x = torch.randn(n, Cin, H, W, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wc = torch.randn(Cout, Cin, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        t0 = bench(lambda: torch.nn.functional.conv2d(x, wc, padding=1))
        rows.append((f"conv3x3 n{n} {H}x{W} {Cin}->{Cout} bn{g.block_n}", ms, fl / ms / 1e9, fl / t0 / 1e9))
    print(f"{'shape':58s} {'ms':>8s} {'ours TF/s':>10s} {'torch TF/s':>10s}")
    for r in rows:
        print(f"{r[0]:58s} {r[1]:8.3f} {r[2]:10.1f} {r[3]:10.1f}")


if __name__ == "__main__":
    main()
