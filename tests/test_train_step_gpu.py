"""The whole training step (BASELINE configs[3], SURVEY.md 8e "training DP" / 8f row 4) on the CUDA library against the
oracle's autograd: loss and EVERY ControlNet parameter gradient of one step (ControlNet -> frozen UNet -> EDM loss + the
one-frame "spatial" pass), plain and bbox models, then that an AdamW step moves the loss the way torch.optim.AdamW does.
Reference: scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1320-1475 as restated by oracle/train.py.
Bar: loss within 5e-3 relative; gradients of all parameters together within 3e-2 relative L2 (bf16 activations and
activation gradients against an fp32 oracle), no large parameter worse than 6e-2."""
import json
import os

import pytest
import torch

from parity_util import oracle_pair, rel_l2, small_cfg

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")


def make_batch(cfg, b=2, h=16, w=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    F = cfg.num_frames
    traj = (torch.rand(b, F, 3, 8 * h, 8 * w, generator=g) > 0.97).float() * 2 - 1
    bbox = (torch.rand(b, F, 3, 8 * h, 8 * w, generator=g) > 0.98).float() * 2 - 1
    return dict(latents=torch.randn(b, F, 4, h, w, generator=g) * 0.18215 * 5, noise=torch.randn(b, F, 4, h, w, generator=g),
                sigmas=torch.tensor([1.3, 0.4, 7.0, 0.05][:b]), image_embeddings=torch.randn(b, 1, cfg.cross_attention_dim, generator=g),
                trajectories=traj, motion_values=torch.tensor([127.0, 90.0, 10.0, 200.0][:b])), bbox


def oracle_step(o_unet, o_cnet, batch, bbox_maps, ran_idx, dev, camera_cond=None, use_spatial=True):
    from oracle.train import training_step
    o_unet.to(dev).requires_grad_(False)
    o_cnet.to(dev).requires_grad_(True)
    for p in o_cnet.parameters():
        p.grad = None

    def cnet(*a, **k):
        if bbox_maps is not None:
            k["controlnet_bbox"] = bbox_maps.to(dev)
        return o_cnet(*a, **k)

    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        kw = {} if camera_cond is None else {"camera_cond": camera_cond.to(dev)}
        out = training_step(o_unet, cnet, ran_idx=ran_idx, use_spatial=use_spatial, **{k: v.to(dev) for k, v in batch.items()}, **kw)
        out["loss"].backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    grads = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for n, p in o_cnet.named_parameters()}
    return out, grads


@pytest.mark.parametrize("variant,h,w,b", [("plain", 16, 24, 2), ("bbox", 16, 24, 2), ("cam", 16, 24, 2), ("plain", 24, 40, 1),
                                           ("plain-nospatial", 16, 24, 2)])
def test_one_step_loss_and_all_gradients(cuda_dev, variant, h, w, b):
    """plain / bbox (second tower through the shared conv_out) / cam (cc_projection) models; the last case has a 3 x 5
    bottom level (odd sizes, partial attention tiles) and batch 1."""
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.train_engine import ControlNetTrainer
    bbox, cam, spatial = variant == "bbox", variant == "cam", not variant.endswith("nospatial")
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=21, bbox=bbox, cam=cam)
    batch, bbox_maps = make_batch(cfg, b=b, h=h, w=w)
    bbox_maps = bbox_maps if bbox else None
    camera = torch.randn(b, cfg.num_frames, 12, generator=torch.Generator().manual_seed(9)) * 0.3 if cam else None
    out, og = oracle_step(o_unet, o_cnet, batch, bbox_maps, 1, cuda_dev, camera, use_spatial=spatial)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev, bbox=bbox, cam=cam)
    tr = ControlNetTrainer(unet, cnet, batch=b, frames=cfg.num_frames, height=h, width=w, use_spatial=spatial)
    loss = tr.forward_backward(ran_idx=1, controlnet_bbox=bbox_maps, camera_cond=camera, **batch)
    tr.buckets.finish()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(out["loss"])) < 5e-3 * abs(float(out["loss"])), (float(loss), float(out["loss"]))
    g = tr.gradients()
    assert set(g) == set(og)
    errs, num, den = {}, 0.0, 0.0
    for k in og:
        d = (g[k].float() - og[k].float())
        num += float(d.pow(2).sum())
        den += float(og[k].float().pow(2).sum())
        n = float(og[k].float().norm())
        errs[k] = (float(d.norm()) / n if n > 0 else float(g[k].float().norm()), n)
    total = (num / den) ** 0.5
    os.makedirs(OUT, exist_ok=True)
    worst = sorted(errs.items(), key=lambda kv: -kv[1][0])[:25]
    with open(os.path.join(OUT, "train_step_parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(variant=variant, hw=[h, w], batch=b, loss=float(loss), oracle_loss=float(out["loss"]), total_rel_l2=total,
                                worst=[(k, round(e, 5), n) for k, (e, n) in worst])) + "\n")
    assert total < 3e-2, (total, worst[:8])
    gnorm = den ** 0.5
    big = {k: e for k, (e, n) in errs.items() if n > 1e-3 * gnorm}
    assert max(big.values()) < 6e-2, sorted(big.items(), key=lambda kv: -kv[1])[:8]
    # parameters autograd leaves untouched (dead cross-attention queries / keys, unused conv_out_2) are exactly zero here too
    for k, (e, n) in errs.items():
        if n == 0:
            assert e == 0, k


def test_full_width_step(cuda_dev):
    """The real channel plan (320 / 640 / 1280 / 1280, heads 5 / 10 / 20 / 20, 682 M trained parameters) on a small clip: the
    tuned tile table, CTA-pair GEMMs, 256-wide wgrad tiles and the tcgen05 attention backward as the full-size step runs
    them, against the fp32 oracle."""
    from posetraj_b200.config import SVDConfig
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.train_engine import ControlNetTrainer
    cfg = SVDConfig(num_frames=2)
    o_unet, o_cnet = oracle_pair(cfg, seed=31)
    batch, _ = make_batch(cfg, b=1, h=16, w=24, seed=8)
    out, og = oracle_step(o_unet, o_cnet, batch, None, 1, cuda_dev)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev)
    del o_unet
    tr = ControlNetTrainer(unet, cnet, batch=1, frames=cfg.num_frames, height=16, width=24)
    loss = tr.forward_backward(ran_idx=1, **batch)
    tr.buckets.finish()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(out["loss"].detach())) < 5e-3 * abs(float(out["loss"].detach()))
    g = tr.gradients()
    num = den = 0.0
    for k in og:
        num += float((g[k].float() - og[k].float()).pow(2).sum())
        den += float(og[k].float().pow(2).sum())
    total = (num / den) ** 0.5
    with open(os.path.join(OUT, "train_step_parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(variant="plain, full width (320-1280 ch), 2 frames", hw=[16, 24], batch=1, loss=float(loss),
                                oracle_loss=float(out["loss"].detach()), total_rel_l2=total, worst=[])) + "\n")
    assert total < 3e-2, total


def test_conditioning_dropout_and_checkpoint_round_trip(cuda_dev, tmp_path):
    """The reference's conditioning dropout (train...cam_concat.py:1365-1385) and save_pretrained -> from_pretrained."""
    from oracle.train import training_step
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.train_engine import ControlNetTrainer
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=23)
    batch, _ = make_batch(cfg, seed=5)
    rp = torch.tensor([0.15, 0.05])            # sample 0: conditioning latent AND embedding dropped (p <= r < 2p); sample 1: embedding
    o_unet.to(cuda_dev).requires_grad_(False)
    o_cnet.to(cuda_dev).requires_grad_(True)
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        out = training_step(o_unet, o_cnet, ran_idx=2, random_p=rp.to(cuda_dev), conditioning_dropout_prob=0.1,
                            **{k: v.to(cuda_dev) for k, v in batch.items()})
        out["loss"].backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev)
    tr = ControlNetTrainer(unet, cnet, batch=2, frames=cfg.num_frames, height=16, width=24, lr=1e-4)
    loss = tr.forward_backward(ran_idx=2, random_p=rp, conditioning_dropout_prob=0.1, **batch)
    tr.buckets.finish()
    assert abs(float(loss) - float(out["loss"].detach())) < 5e-3 * abs(float(out["loss"].detach()))
    g = tr.gradients()
    num = den = 0.0
    for n, p in o_cnet.named_parameters():
        og = p.grad if p.grad is not None else torch.zeros_like(p)
        num += float((g[n].float() - og.float()).pow(2).sum())
        den += float(og.float().pow(2).sum())
    assert (num / den) ** 0.5 < 3e-2
    with pytest.raises(ValueError):
        tr.forward_backward(conditioning_dropout_prob=0.1, **batch)
    # one optimizer step, save in the diffusers layout, load it back into the inference mirror
    tr.set_lr(5e-5)
    tr.optimizer_step()
    path = tr.save_pretrained(str(tmp_path / "controlnet"))
    assert os.path.exists(path)
    again = ControlNetSDVModel.from_pretrained(str(tmp_path), subfolder="controlnet", device=cuda_dev)
    sd, sd2 = tr.state_dict(), again.state_dict()
    assert set(sd) == set(sd2)
    assert all(torch.equal(sd[k].cpu().float(), sd2[k].cpu().float()) for k in sd)
    assert any(not torch.equal(sd[k].cpu().float(), cnet.state_dict()[k].cpu().float()) for k in sd)   # it did move


def test_adamw_step_follows_torch(cuda_dev):
    """Two steps on the same batch: the parameters after our fused AdamW match torch.optim.AdamW driven by the oracle's
    gradients, and the second loss is lower."""
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.train_engine import ControlNetTrainer
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=22)
    batch, _ = make_batch(cfg, seed=3)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev)
    tr = ControlNetTrainer(unet, cnet, batch=2, frames=cfg.num_frames, height=16, width=24, lr=2e-4)
    l0 = float(tr.step(ran_idx=0, **batch))
    out, _ = oracle_step(o_unet, o_cnet, batch, None, 0, cuda_dev)
    opt = torch.optim.AdamW(o_cnet.parameters(), lr=2e-4)
    opt.step()
    sd = tr.state_dict()
    num = den = 0.0
    for n, p in o_cnet.named_parameters():
        num += float((sd[n].float() - p.detach().float()).pow(2).sum())
        den += float(p.detach().float().pow(2).sum())
    assert (num / den) ** 0.5 < 2e-3           # first AdamW step = lr * sign(g): only near-zero gradients may differ
    l1 = float(tr.step(ran_idx=0, **batch))
    assert l1 < l0, (l0, l1)


# -----------------------------------------------------------------------------------------------------------------
# data-parallel step (SURVEY.md 8e "training DP"): world 2 — NCCL on two GPUs (the step replayed as a CUDA graph with the
# bucket all-reduces captured in it), gloo with both ranks on cuda:0 on a 1-GPU box (eager)
# -----------------------------------------------------------------------------------------------------------------
def _dp_worker(rank, port, q, two_gpus):
    try:
        import sys
        import torch.distributed as dist
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        sys.path.insert(0, os.path.join(root, "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
        dev = torch.device("cuda", rank if two_gpus else 0)
        torch.cuda.set_device(dev)
        if two_gpus:
            dist.init_process_group("nccl", rank=rank, world_size=2, device_id=dev)
        else:
            dist.init_process_group("gloo", rank=rank, world_size=2)
        from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
        from posetraj_b200.train_engine import ControlNetTrainer
        cfg = small_cfg()
        unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
        cnet = ControlNetSDVModel.from_random(cfg, dev, seed=3, bbox=True, faithful_zero_init=False)
        batches = []
        for r in range(2):
            b, bm = make_batch(cfg, seed=40 + r)
            b["controlnet_bbox"] = bm
            batches.append({k: v.to(dev) for k, v in b.items()})
        out = {}
        # every rank computes BOTH per-sample-batch gradients alone (no group): the expected average
        solo = ControlNetTrainer(unet, cnet, batch=2, frames=cfg.num_frames, height=16, width=24, group=dist.new_group([rank]))
        want = None
        for r in range(2):
            solo.forward_backward(ran_idx=1, **batches[r])
            solo.buckets.finish()
            g = torch.cat([f.clone() for f in solo.buckets.flat])
            want = g if want is None else want + g
        want = want / 2
        del solo
        tr = ControlNetTrainer(unet, cnet, batch=2, frames=cfg.num_frames, height=16, width=24, lr=1e-4)
        tr.use_cuda_graph = two_gpus
        losses = []
        # step 1 (always eager): the all-reduced buckets hold the SUM of the two ranks' gradients
        loss = tr.forward_backward(ran_idx=1, **batches[rank])
        tr.buckets.finish()
        got = torch.cat([f.clone() for f in tr.buckets.flat]) / 2
        out["grad_err"] = float((got - want).norm() / want.norm())
        losses.append(float(loss))
        tr.optimizer_step()
        for it in range(3):          # graph mode: eager (fills the static inputs), capture + replay, replay
            losses.append(float(tr.step(ran_idx=1, **batches[rank])))
        params = torch.cat([m.reshape(-1) for m in tr.opt.master])
        gathered = [torch.empty_like(params) for _ in range(2)]
        dist.all_gather(gathered, params)
        out["param_diff"] = float((gathered[0] - gathered[1]).abs().max())
        out["losses"] = losses
        q.put((rank, out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))


def test_data_parallel_step_world2():
    import socket
    import torch.multiprocessing as mp
    two_gpus = torch.cuda.device_count() >= 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, port, q, two_gpus)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(2):
            rank, out = q.get(timeout=400)
            assert "error" not in out, out.get("error")
            res[rank] = out
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for r in (0, 1):
        assert res[r]["grad_err"] < 1e-5, res[r]          # all-reduced gradient = mean of the two ranks' gradients
        assert res[r]["param_diff"] == 0.0, res[r]        # both ranks hold identical parameters after 3 steps
        assert all(l == l for l in res[r]["losses"])
