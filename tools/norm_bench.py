"""Micro-benchmark of LayerNorm / GroupNorm at the shapes of one denoise step (CUDA events, L2 flushed)."""
import sys
import torch
sys.path.insert(0, ".")
from posetraj_b200.ops import GroupNorm, LayerNorm

sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def bench(fn, iters=9):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for rows, C in [(80640, 320), (20160, 640), (5040, 1280), (1260, 1280)]:
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(x)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    op = LayerNorm(x, out, g, b)
    us = bench(lambda: op.launch(sp))
    print(f"layernorm {rows}x{C}: {us:.1f} us  {2 * rows * C * 2 / us / 1e3:.0f} GB/s (algorithmic)")
stats = torch.zeros((2 * 28 + 4 * 148 + 64) * 64 + 1024, device="cuda", dtype=torch.float64)
for n_img, H, W, C, per_stat_frames, halo in [(28, 40, 72, 320, 1, True), (28, 40, 72, 320, 14, False), (28, 20, 36, 640, 1, True),
                                              (28, 20, 36, 640, 14, False), (28, 10, 18, 1280, 1, True), (28, 5, 9, 1280, 1, True),
                                              (28, 5, 9, 1280, 14, False), (28, 40, 72, 640, 1, True)]:
    x = torch.randn(n_img * H * W, C, device="cuda").to(torch.bfloat16)
    rows_out = n_img * (H + 1) * (W + 1) if halo else n_img * H * W
    out = torch.empty(rows_out, C, device="cuda", dtype=torch.bfloat16)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for silu in (True, False):
        op = GroupNorm(x, out, g, b, stats, rows_per_stat=per_stat_frames * H * W, eps=1e-5, silu=silu, halo=(H, W) if halo else None)
        us = bench(lambda: op.launch(sp))
        print(f"groupnorm {n_img}x{H}x{W}x{C} frames/stat={per_stat_frames} halo={halo} silu={silu}: {us:.1f} us  {2 * x.numel() * 2 / us / 1e3:.0f} GB/s (algorithmic)")
