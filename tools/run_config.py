"""Runs N denoise steps of a BASELINE.json config shape on one GPU and reports ms/step (CUDA-graph replay).
Usage: python tools/run_config.py FRAMES LAT_H LAT_W [cam] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.config import SVDConfig
from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
from posetraj_b200.roofline import step_flops
from posetraj_b200.trajectory import rasterize_tracks

F, h, w = (int(v) for v in sys.argv[1:4])
cam = "cam" in sys.argv
steps = int(sys.argv[-1]) if sys.argv[-1].isdigit() and len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
cfg = SVDConfig(num_frames=F)
unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
cnet = ControlNetSDVModel.from_random(cfg, dev, seed=0, faithful_zero_init=False, cam=cam)
pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
g = torch.Generator().manual_seed(1234)
img = torch.randn(1, 4, h, w, generator=g)
emb = torch.randn(1, 1, cfg.cross_attention_dim, generator=g)
tracks = [[[20 + 9 * k, 30 + 5 * k] for k in range(F)]]
cond = rasterize_tracks(tracks, F, h * 8, w * 8, dev)
kw = dict(height=h * 8, width=w * 8, num_frames=F, num_inference_steps=steps, output_type="latent",
          latents=torch.randn(1, F, 4, h, w, generator=g), image_embeddings=torch.cat([torch.zeros_like(emb), emb]),
          image_latents=torch.cat([torch.zeros_like(img), img]))
if cam:
    kw["camera_cond"] = torch.randn(F, 12, generator=g) * 0.1
out = pipe(None, cond, **kw).frames
torch.cuda.synchronize()
eng = pipe.engine_for(F, h, w, (h * 8, w * 8))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.reset()
e0.record()
for _ in range(steps):
    eng.graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fl, _ = step_flops(cfg, frames=F, h=h, w=w, cam=cam, essential=True)
print(f"frames={F} latent={h}x{w} cam={cam}: finite={bool(torch.isfinite(out).all())} {ms:.2f} ms/step "
      f"{fl / 1e12:.1f} TFLOP/step -> {fl / ms / 1e9:.0f} TFLOP/s; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; "
      f"{eng.launches_per_step} launches/step")
