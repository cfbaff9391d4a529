"""The "library bar": the oracle's wiring run as plain torch eager bf16 on the same B200 (cuDNN / cuBLAS / SDPA
kernels, NCHW, no fusion) — what the reference would do on this GPU if diffusers were installed, since the reference
ships no Blackwell kernel of its own.  Test infrastructure (imports oracle/): never part of the product path.
Lives under tests/ because only test infrastructure may import oracle/.  Usage: python tests/torch_eager_baseline.py [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.models import build_models
from oracle.pipeline import denoise_step, make_inputs
from oracle.scheduler import EulerKarrasOracle

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
t0 = time.time()
with torch.device("meta"):
    pass
unet, cnet = build_models(seed=0, randomize_zero_convs=True)
unet = unet.to(dev, torch.bfloat16)
cnet = cnet.to(dev, torch.bfloat16)
inp = {k: (v.to(dev, torch.bfloat16) if torch.is_tensor(v) else v) for k, v in make_inputs().items()}
cond = torch.full((2, 14, 3, 320, 576), -1.0, device=dev, dtype=torch.bfloat16)
sched = EulerKarrasOracle()
sched.set_timesteps(25, device=dev)
lat = inp["latents"]
print(f"setup {time.time() - t0:.1f} s", flush=True)
for fmt in ("contiguous", "channels_last"):
    if fmt == "channels_last":
        for m in list(unet.modules()) + list(cnet.modules()):
            if isinstance(m, torch.nn.Conv2d):
                m.to(memory_format=torch.channels_last)
    times = []
    for i in range(steps + 2):
        sched._step_index = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = denoise_step(unet, cnet, sched, lat, 0, sched.timesteps[0], inp["image_latents"], inp["image_embeddings"], cond,
                           inp["added_time_ids"], inp["guidance"])
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    print(f"torch eager bf16 ({fmt}): {ms:.1f} ms/step -> {1000 / ms:.2f} steps/s (finite={bool(torch.isfinite(out.float()).all())})", flush=True)
