"""Sweeps block_n x {single CTA, CTA pair} for the GEMM shapes of one denoise step (CUDA events, L2 flushed)."""
import math
import sys

import torch

sys.path.insert(0, ".")
from posetraj_b200.ops import Gemm, conv3x3_taps  # noqa: E402


def bench(fn, iters=7):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(2):
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
    print("kind rows N K taps geglu res | " + " ".join(f"{'s' if not pr else 'p'}{bn}" for pr in (False, True) for bn in (64, 96, 128, 160, 192, 224, 256)))
    shapes = []
    # (kind, geometry, N, K, geglu, res)
    for (M, N, K, geglu, res) in [(80640, 320, 320, False, True), (80640, 960, 320, False, False), (80640, 1280, 320, True, False),
                                  (80640, 320, 1280, False, True), (20160, 640, 640, False, True), (20160, 1920, 640, False, False),
                                  (20160, 2560, 640, True, False), (20160, 640, 2560, False, True), (5040, 1280, 1280, False, True),
                                  (5040, 3840, 1280, False, False), (5040, 5120, 1280, True, False), (5040, 1280, 5120, False, True),
                                  (1260, 1280, 1280, False, True), (1260, 5120, 1280, True, False), (1260, 1280, 5120, False, True)]:
        shapes.append(("lin", (M,), N, K, geglu, res))
    for (n, H, W, Cin, Cout) in [(28, 40, 72, 320, 320), (28, 20, 36, 640, 640), (28, 10, 18, 1280, 1280), (28, 5, 9, 1280, 1280),
                                 (28, 40, 72, 640, 320), (28, 20, 36, 1280, 640), (28, 10, 18, 2560, 1280)]:
        shapes.append(("conv", (n, H, W), Cout, Cin, False, False))
    for (B, Fr, HW, Cc) in [(2, 14, 2880, 320), (2, 14, 720, 640), (2, 14, 180, 1280), (2, 14, 45, 1280)]:
        shapes.append(("tconv", (B, Fr, HW), Cc, Cc, False, True))
    for kind, geo, N, K, geglu, res in shapes:
        if kind == "lin":
            M = geo[0]
            a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            w = (torch.randn((2 * N if geglu else N), K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
            out_rows, kw, taps, fl = M, {}, 1, 2.0 * M * w.shape[0] * K
        elif kind == "conv":
            n, H, W = geo
            a = torch.randn(n * (H + 1) * (W + 1), K, device="cuda").to(torch.bfloat16)
            w = (torch.randn(N, 9 * K, device="cuda") / math.sqrt(9 * K)).to(torch.bfloat16)
            out_rows, kw, taps, fl = n * H * W, dict(taps=conv3x3_taps(W), halo=(H, W)), 9, 2.0 * n * H * W * N * 9 * K
        else:
            B, Fr, HW = geo
            a = torch.randn(B * Fr * HW, K, device="cuda").to(torch.bfloat16)
            w = (torch.randn(N, 3 * K, device="cuda") / math.sqrt(3 * K)).to(torch.bfloat16)
            out_rows, kw, taps, fl = B * Fr * HW, dict(taps=(-HW, 0, HW), batches=B), 3, 2.0 * B * Fr * HW * N * 3 * K
        bias = torch.randn(w.shape[0], device="cuda")
        out = torch.empty(out_rows, N, device="cuda", dtype=torch.bfloat16)
        r = torch.randn(out_rows, N, device="cuda").to(torch.bfloat16) if res else None
        cells = []
        best = (1e9, None)
        for pair in (False, True):
            for bn in (64, 96, 128, 160, 192, 224, 256):
                if geglu and bn % 64:
                    cells.append("   -  ")
                    continue
                per = bn // 2 if geglu else bn
                if per > max(64, ((N + 31) // 32) * 32):
                    cells.append("   -  ")
                    continue
                try:
                    g = Gemm(a, w, out, bias=bias, geglu=geglu, res1=r, block_n=bn, cta_pair=pair, **kw)
                    us = bench(lambda: g.launch(sp)) * 1e3
                except Exception as e:  # noqa: BLE001
                    cells.append(" err  ")
                    continue
                cells.append(f"{us:6.1f}")
                if us < best[0]:
                    best = (us, f"{'p' if pair else 's'}{bn}")
        print(f"{kind} {a.shape[0]} {N} {K} {taps} {int(geglu)} {int(res)} | " + " ".join(cells) + f" | best {best[1]} {best[0]:.1f}us {fl / best[0] / 1e6:.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
