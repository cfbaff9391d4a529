"""Micro-benchmark of LayerNorm and temporal attention at the shapes of one denoise step (CUDA events, L2 flushed).
The kernel variants are chosen by environment variables read once per process (PT_LN_PACKED, PT_LN_FULL, PT_TATTN_STAGED),
so tools/r3_ln.sh runs this once per variant."""
import os
import sys
import torch
sys.path.insert(0, ".")
from posetraj_b200.ops import AttnTemporal, LayerNorm

sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
print({k: os.environ.get(k) for k in ("PT_LN_PACKED", "PT_LN_FULL", "PT_TATTN_STAGED")})


def bench(fn, iters=11):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for rows, C, hw in [(80640, 320, 2880), (20160, 640, 720), (5040, 1280, 180), (1260, 1280, 45)]:
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(x)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    op = LayerNorm(x, out, g, b)
    us = bench(lambda: op.launch(sp))
    print(f"layernorm {rows}x{C}: {us:.1f} us  {2 * rows * C * 2 / us / 1e3:.0f} GB/s (algorithmic)")
    pos = torch.randn(14, C, device="cuda")
    mix = torch.empty_like(x)
    op2 = LayerNorm(x, out, g, b, addvec=pos, hw=hw, frames=14, sum_out=mix)
    us = bench(lambda: op2.launch(sp))
    print(f"layernorm+pos {rows}x{C}: {us:.1f} us  {3 * rows * C * 2 / us / 1e3:.0f} GB/s (algorithmic)")
for B, Fr, HW, heads in [(2, 14, 2880, 5), (2, 14, 720, 10), (2, 14, 180, 20), (2, 14, 45, 20), (2, 25, 9216, 5)]:
    C = heads * 64
    qkv = torch.randn(B * Fr * HW, 3 * C, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * Fr * HW, C, device="cuda", dtype=torch.bfloat16)
    op = AttnTemporal(qkv, out, batch=B, frames=Fr, hw=HW, heads=heads)
    us = bench(lambda: op.launch(sp))
    print(f"attn_temporal B{B} F{Fr} HW{HW} heads{heads}: {us:.1f} us  {4 * B * Fr * HW * C * 2 / us / 1e3:.0f} GB/s (algorithmic)")
