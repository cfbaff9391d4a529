"""One eager denoise step of the full-size hot path between cudaProfilerStart/Stop — the target of the ncu recipes
(`ncu --profile-from-start off ...`).  Usage: python tools/profile_step.py [frames h w]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posetraj_b200.config import SVDConfig
from posetraj_b200.engine import NetPlan
from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet

F, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (14, 40, 72)
dev = torch.device("cuda:0")
cfg = SVDConfig()
unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
cnet = ControlNetSDVModel.from_random(cfg, dev, seed=0, faithful_zero_init=False)
pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
pipe.use_cuda_graph = False
g = torch.Generator().manual_seed(1234)
img = torch.randn(1, 4, h, w, generator=g)
emb = torch.randn(1, 1, cfg.cross_attention_dim, generator=g)
cond = torch.full((F, 3, h * 8, w * 8), -1.0)
cond[:, 0, 100:140, 100:300] = 1.0
out = pipe(None, cond, height=h * 8, width=w * 8, num_frames=F, num_inference_steps=2,
           latents=torch.randn(1, F, 4, h, w, generator=g), output_type="latent",
           image_embeddings=torch.cat([torch.zeros_like(emb), emb]), image_latents=torch.cat([torch.zeros_like(img), img]))
eng = pipe.engine_for(F, h, w, (h * 8, w * 8))
eng.reset()
torch.cuda.synchronize()
# launch-ordered op list (GroupNorm = 2 kernels) so the ncu launch list can be joined with layer names offline
import json
oplist = []
for op in eng.step_ops:
    d = {"name": op.name, "kind": op.kind, "flops": op.alg_flops, "bytes": op.alg_bytes,
         "kernels": 1}
    if op.kind == "gemm":
        a = op.args
        d.update(block_n=a.block_n, n_out=a.n_out, rows=a.rows_per_batch * a.batches, taps=a.num_taps,
                 k=(a.k0_chunks + a.k1_chunks) * 64, geglu=a.geglu)
    oplist.append(d)
os.makedirs("gpurun_out", exist_ok=True)
with open(os.environ.get("PT_OPLIST", "gpurun_out/oplist.json"), "w") as f:
    json.dump(oplist, f)
torch.cuda.profiler.start()
NetPlan.run(eng.step_ops, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step:", eng.launches_per_step, "launches; finite:", bool(torch.isfinite(eng.latents).all()))
