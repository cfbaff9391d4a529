"""One video frame-sharded over the ranks of a torchrun job (SURVEY.md §8e "Frames", BASELINE.json configs[4]).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           tools/run_sharded.py FRAMES LAT_H LAT_W [steps]
Prints ms/step as the max over ranks (device-timed)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.config import SVDConfig
from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
from posetraj_b200.roofline import step_flops
from posetraj_b200.trajectory import rasterize_tracks

F, h, w = (int(v) for v in sys.argv[1:4])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = SVDConfig(num_frames=F)
unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
cnet = ControlNetSDVModel.from_random(cfg, dev, seed=0, faithful_zero_init=False)
pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
pipe.enable_frame_sharding(rank, world)
g = torch.Generator().manual_seed(1234)
img = torch.randn(1, 4, h, w, generator=g)
emb = torch.randn(1, 1, cfg.cross_attention_dim, generator=g)
cond = rasterize_tracks([[[20 + 9 * k, 30 + 5 * k] for k in range(F)]], F, h * 8, w * 8, dev)
kw = dict(height=h * 8, width=w * 8, num_frames=F, num_inference_steps=3, output_type="latent",
          latents=torch.randn(1, F, 4, h, w, generator=g), image_embeddings=torch.cat([torch.zeros_like(emb), emb]),
          image_latents=torch.cat([torch.zeros_like(img), img]))
out = pipe(None, cond, **kw).frames
eng = pipe.engine_for(F, h, w, (h * 8, w * 8))
if "eager" in sys.argv:
    eng.graph = None
eng.reset()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.step()
e1.record()
dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    fl, _ = step_flops(cfg, frames=F, h=h, w=w, essential=True)
    ms = float(t.item())
    print(f"frame-sharded x{world} ({'graph' if eng.graph is not None else 'eager'}): frames={F} latent={h}x{w} finite={bool(torch.isfinite(out).all())} {ms:.2f} ms/step "
          f"({fl / ms / 1e9:.0f} TFLOP/s aggregate), {eng.launches_per_step} launches + {eng.collectives_per_step} collectives per step, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB/rank")
# captured NCCL graphs keep communicator work alive: drop them, sync, and leave without the (slow) collective teardown
eng.graph = None
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
