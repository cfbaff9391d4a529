"""Per-parameter gradient errors of one training step against the oracle (small config), for debugging on a GPU box."""
import sys, os, json
sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from parity_util import oracle_pair, small_cfg
from test_train_step_gpu import make_batch, oracle_step
from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from posetraj_b200.train_engine import ControlNetTrainer

bbox = len(sys.argv) > 1 and sys.argv[1] == "bbox"
use_spatial = not (len(sys.argv) > 2 and sys.argv[2] == "nospatial")
dev = torch.device("cuda", 0)
cfg = small_cfg()
o_unet, o_cnet = oracle_pair(cfg, seed=21, bbox=bbox)
batch, bm = make_batch(cfg)
bm = bm if bbox else None
if use_spatial:
    out, og = oracle_step(o_unet, o_cnet, batch, bm, 1, dev)
else:
    from oracle.train import training_step
    o_unet.to(dev).requires_grad_(False); o_cnet.to(dev).requires_grad_(True)
    out = training_step(o_unet, o_cnet, ran_idx=1, use_spatial=False, **{k: v.to(dev) for k, v in batch.items()})
    out["loss"].backward()
    og = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for n, p in o_cnet.named_parameters()}
unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), dev)
cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), dev, bbox=bbox)
tr = ControlNetTrainer(unet, cnet, batch=2, frames=cfg.num_frames, height=16, width=24, use_spatial=use_spatial)
loss = tr.forward_backward(ran_idx=1, controlnet_bbox=bm, **batch)
tr.buckets.finish()
torch.cuda.synchronize()
print("loss", float(loss), "oracle", float(out["loss"]))
g = tr.gradients()
rows = []
num = den = 0.0
for k in og:
    d = g[k].float() - og[k].float()
    n = float(og[k].float().norm())
    num += float(d.pow(2).sum()); den += n * n
    rows.append((float(d.norm()) / n if n > 0 else float(g[k].float().norm()), n, k))
print("total rel l2", (num / den) ** 0.5, "grad norm", den ** 0.5)
for e, n, k in sorted(rows, reverse=True)[:60]:
    print(f"{e:10.4f} {n:12.5g} {k}")
print("...")
for e, n, k in sorted(rows)[:10]:
    print(f"{e:10.4f} {n:12.5g} {k}")
