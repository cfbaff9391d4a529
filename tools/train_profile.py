"""Where the time of the configs[3] training step goes: every backward helper of posetraj_b200.training wrapped with a
device-synchronised timer (inclusive times; `wgrad` contains `transpose`), plus the forward op lists by kernel class."""
import json, os, sys, time, collections
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from posetraj_b200 import training, train_engine, ops
from posetraj_b200.config import SVDConfig
from posetraj_b200.engine import NetPlan
from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from posetraj_b200.train_engine import ControlNetTrainer

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
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

acc = collections.defaultdict(lambda: [0.0, 0])


def wrap(mod, name, label=None):
    fn = getattr(mod, name)
    label = label or name

    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label][0] += (time.perf_counter() - t0) * 1e3
        acc[label][1] += 1
        return r
    setattr(mod, name, w)


for n in ("linear_dgrad", "wgrad", "transpose", "groupnorm_backward", "layernorm_backward", "attention_spatial_backward",
          "attention_temporal_backward", "colsum_grouped", "to_halo", "dilate2x", "zero_halo", "geglu_backward", "silu_backward",
          "dgrad_weight", "upsample_backward", "edm_loss_into"):
    wrap(training, n)
wrap(train_engine.Tape, "add", "Tape.add (Axpy accumulate)")
wrap(train_engine.Tape, "mark", "Tape.mark")
orig_run = NetPlan.run


def run(op_list, stream_ptr=None):
    by = collections.defaultdict(float)
    for op in op_list:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        op.launch(stream_ptr if stream_ptr is not None else torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        k = "fwd:" + getattr(op, "kind", type(op).__name__)
        acc[k][0] += (time.perf_counter() - t0) * 1e3
        acc[k][1] += 1


NetPlan.run = staticmethod(run)
wrap(tr, "optimizer_step", "optimizer_step (all-reduce wait + AdamW + refresh)")
torch.cuda.synchronize()
t0 = time.perf_counter()
tr.forward_backward(ran_idx=3, **batch)
torch.cuda.synchronize()
t_fb = (time.perf_counter() - t0) * 1e3
tr.optimizer_step()
rows = sorted(acc.items(), key=lambda kv: -kv[1][0])
print(f"forward_backward with per-call synchronisation: {t_fb:.1f} ms")
print("| helper | calls | ms (inclusive, synchronised) |\n|---|---|---|")
for k, (ms, n) in rows:
    print(f"| {k} | {n} | {ms:.2f} |")
