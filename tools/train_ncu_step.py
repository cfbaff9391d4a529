"""One eager configs[3] training step (forward + backward, no optimizer) between cudaProfilerStart/Stop — target of
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file X python tools/train_ncu_step.py
`python tools/train_ncu_step.py summarize X [out.md]` aggregates the launch list by kernel name."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))

if len(sys.argv) > 1 and sys.argv[1] == "summarize":
    import csv, collections
    with open(sys.argv[2]) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r["Kernel Name"].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    out = [f"device time of one training step (forward + backward; ncu, serialised): {tot:.1f} ms over {len(rows)} launches\n",
           "| kernel | launches | ms | share | avg us |", "|---|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {ms:.2f} | {100 * ms / tot:.1f}% | {1e3 * ms / n:.1f} |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(text + "\n")
    sys.exit(0)

import torch
from posetraj_b200.config import SVDConfig
from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from posetraj_b200.train_engine import ControlNetTrainer

dev = torch.device("cuda", 0)
cfg = SVDConfig()
B, Fr, H, W = 2, 14, 40, 72
unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
cnet = ControlNetSDVModel.from_random(cfg, dev, seed=5, bbox=True, faithful_zero_init=False)
tr = ControlNetTrainer(unet, cnet, batch=B, frames=Fr, height=H, width=W)
g = torch.Generator(device=dev).manual_seed(11)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
batch = dict(latents=rn(B, Fr, 4, H, W) * 0.9, noise=rn(B, Fr, 4, H, W), sigmas=torch.tensor([1.3, 0.4], device=dev),
             image_embeddings=rn(B, 1, cfg.cross_attention_dim),
             trajectories=(torch.rand(B, Fr, 3, 8 * H, 8 * W, device=dev, generator=g) > 0.97).float() * 2 - 1,
             motion_values=torch.tensor([127.0, 90.0], device=dev),
             controlnet_bbox=(torch.rand(B, Fr, 3, 8 * H, 8 * W, device=dev, generator=g) > 0.98).float() * 2 - 1)
tr.step(ran_idx=3, **batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.forward_backward(ran_idx=3, **batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one forward_backward")
