"""Full-size VAE timing on one B200: decode 14 frames at 320x576 (latent 40x72) and encode one 320x576 image with the
SVD-shaped random-init VAE (97.7 M parameters); CUDA events, median of 5; per-kernel-class split of the decode."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from posetraj_b200.vae import AutoencoderKLTemporalDecoder, VaeConfig  # noqa: E402

dev = torch.device("cuda", 0)
frames, h, w = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (14, 40, 72)
vae = AutoencoderKLTemporalDecoder.from_random(VaeConfig(), dev, seed=0)
z = torch.randn(frames, 4, h, w, device=dev) / 0.18215
img = torch.rand(1, 3, 8 * h, 8 * w, device=dev) * 2 - 1


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


t0 = time.time()
out = vae.decode(z, num_frames=frames).sample
torch.cuda.synchronize()
first = time.time() - t0
assert torch.isfinite(out).all()
dec_ms = timed(lambda: vae.decode(z, num_frames=frames))
enc_ms = timed(lambda: vae.encode(img))
plan = vae._dec[(1, frames, h, w)]
flops = sum(getattr(op, "alg_flops", 0.0) for op in plan.ops)
# per-class device time of one decode (events around every launch)
sp = torch.cuda.current_stream().cuda_stream
cls = {}
for op in plan.ops:
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); op.launch(sp); e.record(); torch.cuda.synchronize()
    k = getattr(op, "kind", "misc")
    c = cls.setdefault(k, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
    c["ms"] += s.elapsed_time(e); c["launches"] += 1
    c["flops"] += getattr(op, "alg_flops", 0.0); c["bytes"] += getattr(op, "alg_bytes", 0.0)
for k, c in cls.items():
    c["tflops"] = round(c["flops"] / c["ms"] / 1e9, 1) if c["flops"] else None
    c["gbs"] = round(c["bytes"] / c["ms"] / 1e6, 1) if c["bytes"] else None
    c["ms"] = round(c["ms"], 3); del c["flops"]; del c["bytes"]
print(json.dumps({"what": "vae", "frames": frames, "latent": [h, w], "decode_ms": round(dec_ms, 2), "encode_ms": round(enc_ms, 2),
                  "decode_tflop": round(flops / 1e12, 2), "decode_tflops": round(flops / dec_ms / 1e9, 1),
                  "launches": len(plan.ops), "first_call_s": round(first, 2), "pool_GiB": round(plan.pool.bytes / 2**30, 2),
                  "mem_GiB": round(torch.cuda.max_memory_allocated() / 2**30, 2), "classes": cls}))
