"""Model-level parity AT THE BASELINE SHAPE (BASELINE.json configs[1]/[2]/[4]): the full-width SVD-shaped networks
(320/640/1280/1280 channels, heads 5/10/20/20, 1524.6 M + ControlNetSDV with the camera branch and the bbox tower),
CFG pair x 14 frames x 40x72 latent — the configuration the benchmark is quoted on — against the oracle executed on
the same GPU in fp32 with TF32 off (cuDNN / cuBLAS / SDPA fp32: the reference's default inference precision).

This is where the tuned tile table (ops.TUNED, CTA-pair tiles, block_n 160..256), the 5/10/20-head attention launches
and the co-resident GroupNorm grid at 80 640 rows run together; the small-config tests never reach those code paths.

Tolerances (north_star): per-step noise prediction relative L2 <= 1e-2 in bf16.  ControlNet residuals: 2e-2 (they
are 13 intermediate tensors up to 27 blocks deep; torch's own bf16 run of the oracle is recorded beside ours).
Measured values are appended to gpurun_out/parity_fullshape.jsonl and summarised in profiles/r2_parity_fullshape.md.
"""
import json
import math
import os

import pytest
import torch

from parity_util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-2
TOL_RES = 2e-2


def _record(name, value):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_fullshape.jsonl", "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


def build_full_oracle(dev, cam=True, bbox=True, seed=0):
    """The oracle pair at full width, constructed directly on the GPU (torch default inits under manual_seed), zero
    convs re-drawn N(0, 1/fan_in) so that the ControlNet branch is exercised (weight set W1), matrices
    rounded to bf16 values so both sides hold identical weights."""
    from oracle.models import build_models
    torch.manual_seed(seed)
    with torch.device(dev):
        unet, cnet = build_models(seed=seed, cam=cam, bbox=bbox, randomize_zero_convs=False)
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    with torch.no_grad():
        convs = list(cnet.controlnet_down_blocks) + [cnet.controlnet_mid_block, cnet.controlnet_cond_embedding.conv_out]
        for conv in convs:
            fan_in = conv.weight[0].numel()
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g, device=dev) / fan_in ** 0.5)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=g, device=dev) * 0.02)
        if cam:
            # a camera projection that is not the identity, so the 12 camera columns matter
            pj = cnet.controlnet_cond_embedding.cc_projection
            pj.weight.add_(torch.randn(pj.weight.shape, generator=g, device=dev) * 0.05)
        for m in (unet, cnet):
            for p in m.parameters():
                if p.dim() > 1:
                    p.copy_(p.to(torch.bfloat16).float())
    return unet, cnet


def full_inputs(frames, h, w, dev, seed=1234):
    from oracle.pipeline import make_inputs
    inp = make_inputs(num_frames=frames, h=h, w=w, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    H, W = h * 8, w * 8

    def maps(shift):
        cond = torch.full((frames, 3, H, W), -1.0)
        for f in range(frames):   # a trajectory-drawing-like pattern: a red bar and a green blob moving over frames
            y0, x0 = (40 + 11 * f + shift) % (H - 20), (60 + 23 * f + 2 * shift) % (W - 80)
            cond[f, 0, y0:y0 + 7, x0:x0 + 60] = 1.0
            cond[f, 1, y0 + 3:y0 + 10, x0 + 54:x0 + 61] = 1.0
        return torch.cat([cond[None]] * 2)

    inp["controlnet_condition"] = maps(0)
    inp["controlnet_bbox"] = maps(37)
    inp["camera_cond"] = torch.cat([torch.randn(1, frames, 12, generator=g) * 0.1] * 2)
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}


@pytest.fixture(scope="module")
def full_setup(cuda_dev):
    from posetraj_b200.config import SVDConfig
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    cfg = SVDConfig()
    o_unet, o_cnet = build_full_oracle(cuda_dev)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev, cam=True, bbox=True)
    return cfg, o_unet, o_cnet, unet, cnet


def model_input(inp, sigma):
    x = torch.cat([inp["latents"]] * 2) / (sigma ** 2 + 1) ** 0.5
    return torch.cat([x, inp["image_latents"]], dim=2)


def _one_step(full_setup, inp, sigma, variant, scale=0.8):
    cfg, o_unet, o_cnet, unet, cnet = full_setup
    x = model_input(inp, sigma)
    t = torch.tensor(0.25 * math.log(sigma), device=x.device)
    kw = {}
    if variant == "cam":
        kw["camera_cond"] = inp["camera_cond"]
    if variant == "bbox":
        kw["controlnet_bbox"] = inp["controlnet_bbox"]
    with torch.no_grad():
        o_down, o_mid = o_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"],
                               controlnet_cond=inp["controlnet_condition"], conditioning_scale=scale, **kw)
        o_pred = o_unet(x, t, inp["image_embeddings"], down_block_additional_residuals=o_down,
                        mid_block_additional_residual=o_mid, added_time_ids=inp["added_time_ids"])
    down, mid = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"],
                     conditioning_scale=scale, return_dict=False, **kw)
    torch.cuda.synchronize()
    errs = [rel_l2(a, b) for a, b in zip(down + [mid], o_down + [o_mid])]
    pred = unet(x, t, inp["image_embeddings"], down_block_additional_residuals=down, mid_block_additional_residual=mid,
                added_time_ids=inp["added_time_ids"], return_dict=False)[0]
    torch.cuda.synchronize()
    assert pred.shape == o_pred.shape and torch.isfinite(pred).all()
    return errs, rel_l2(pred, o_pred), (o_down, o_mid, o_pred)


@pytest.mark.parametrize("sigma,variant", [(700.0, "plain"), (10.0, "cam"), (0.05, "bbox")])
def test_full_shape_single_step(full_setup, cuda_dev, sigma, variant):
    """configs[1] (plain), configs[2] (camera branch) and the bbox tower of configs[3], one step each, at three noise
    levels spanning the Karras schedule."""
    inp = full_inputs(14, 40, 72, cuda_dev)
    errs, e_pred, _ = _one_step(full_setup, inp, sigma, variant)
    _record(f"full 14x40x72 sigma={sigma} {variant}: 13 residuals", errs)
    _record(f"full 14x40x72 sigma={sigma} {variant}: noise_pred", e_pred)
    assert max(errs) < TOL_RES, errs
    assert e_pred < TOL, e_pred


def test_full_shape_variants_differ(full_setup, cuda_dev):
    """The camera / bbox inputs must actually change the residuals at full shape (guards against a test that
    passes because a branch is dead)."""
    cfg, o_unet, o_cnet, unet, cnet = full_setup
    inp = full_inputs(14, 40, 72, cuda_dev)
    x = model_input(inp, 10.0)
    t = torch.tensor(0.25 * math.log(10.0), device=cuda_dev)
    base = [r.clone() for r in cnet(x, t, inp["image_embeddings"], inp["added_time_ids"],
                                    controlnet_cond=inp["controlnet_condition"], return_dict=False)[0]]
    cam = [r.clone() for r in cnet(x, t, inp["image_embeddings"], inp["added_time_ids"],
                                   controlnet_cond=inp["controlnet_condition"], camera_cond=inp["camera_cond"],
                                   return_dict=False)[0]]
    bbox = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"],
                controlnet_bbox=inp["controlnet_bbox"], return_dict=False)[0]
    torch.cuda.synchronize()
    assert rel_l2(cam[0], base[0]) > 1e-3 and rel_l2(bbox[0], base[0]) > 1e-3


def test_full_shape_torch_bf16_context(full_setup, cuda_dev):
    """Context for the tolerance at full shape: torch's own bf16 run of the oracle (cuDNN / cuBLAS / SDPA bf16) against
    the same fp32 oracle output — recorded, and our error must not be worse than 1.5x it (+1e-3)."""
    import copy
    cfg, o_unet, o_cnet, unet, cnet = full_setup
    inp = full_inputs(14, 40, 72, cuda_dev)
    sigma = 10.0
    errs, e_ours, (o_down, o_mid, want) = _one_step(full_setup, inp, sigma, "plain", scale=1.0)
    x = model_input(inp, sigma)
    t = torch.tensor(0.25 * math.log(sigma), device=cuda_dev)
    with torch.no_grad():
        bu = copy.deepcopy(o_unet).to(torch.bfloat16)
        bc = copy.deepcopy(o_cnet).to(torch.bfloat16)
        d16 = {k: (v.to(torch.bfloat16) if torch.is_tensor(v) else v) for k, v in inp.items()}
        b_down, b_mid = bc(x.to(torch.bfloat16), t, d16["image_embeddings"], d16["added_time_ids"],
                           controlnet_cond=d16["controlnet_condition"])
        torch_bf16 = bu(x.to(torch.bfloat16), t, d16["image_embeddings"], down_block_additional_residuals=b_down,
                        mid_block_additional_residual=b_mid, added_time_ids=d16["added_time_ids"])
        e_torch_res = [rel_l2(a, b) for a, b in zip(b_down + [b_mid], o_down + [o_mid])]
    e_torch = rel_l2(torch_bf16, want)
    del bu, bc
    _record("full 14x40x72 sigma=10 plain: noise_pred [ours, torch-bf16]", [e_ours, e_torch])
    _record("full 14x40x72 sigma=10 plain: residuals ours", errs)
    _record("full 14x40x72 sigma=10 plain: residuals torch-bf16", e_torch_res)
    assert e_ours < TOL
    assert e_ours < 1.5 * e_torch + 1e-3


def test_full_shape_two_step_latents(full_setup, cuda_dev):
    """The fused loop at full shape (CUDA-graph replay from step 1, CFG + Euler kernel, device-side sigma table) against
    the oracle's loop, on a complete 2-step Karras schedule (sigma 700 -> 0.002 -> 0): the final latents are then
    essentially the networks' denoised prediction, so the comparison is as sharp as the single-step one (with a long
    schedule the first steps barely move the sigma-700 latents and any model error drowns).  conditioning_scale != 1
    goes through the captured graph, then a second load with another scale replays the SAME graph."""
    from oracle.pipeline import denoise
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    cfg, o_unet, o_cnet, unet, cnet = full_setup
    inp = full_inputs(14, 40, 72, cuda_dev, seed=99)
    steps = 2
    pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
    pipe.scheduler.set_timesteps(steps, device=cuda_dev)
    eng = pipe.engine_for(14, 40, 72, (320, 576))
    res = {}
    for scale in (0.7, 1.6):
        want = denoise(o_unet, o_cnet, inp["latents"], inp["image_latents"], inp["image_embeddings"],
                       inp["controlnet_condition"], inp["added_time_ids"], inp["guidance"], num_inference_steps=steps,
                       cond_scale=scale)
        eng.load(latents=inp["latents"], image_latents=inp["image_latents"], image_embeddings=inp["image_embeddings"],
                 added_time_ids=inp["added_time_ids"], guidance=inp["guidance"], sigmas=pipe.scheduler.sigmas,
                 controlnet_condition=inp["controlnet_condition"], cond_scale=scale)
        if eng.graph is None:
            eng.step(use_graph=False)      # first call: one eager step, then capture (as __call__ does)
            eng.capture()
            eng.step(use_graph=True)
        else:
            eng.step(use_graph=True)       # second call: both steps replay the graph captured at the other scale
            eng.step(use_graph=True)
        torch.cuda.synchronize()
        got = eng.latents.view(1, 14, 4, 40, 72).clone()
        e = rel_l2(got, want)
        _record(f"full 14x40x72: latents after a complete 2-step schedule, cond_scale {scale} (graph replay)", e)
        assert e < 2 * TOL, (scale, e)
        res[scale] = want
    d = rel_l2(res[1.6], res[0.7])
    _record("full 14x40x72: distance between the cond_scale 0.7 and 1.6 results (oracle)", d)
    assert d > 4 * TOL, d   # far more than the bound above: replaying a graph-baked 0.7 would have failed it


def test_full_shape_config5_one_step(full_setup, cuda_dev):
    """configs[4] shape: 25 frames, 72x128 latent (576x1024 px), one step, single GPU (the unsharded plan the
    frame-sharded run is compared with)."""
    inp = full_inputs(25, 72, 128, cuda_dev, seed=5)
    errs, e_pred, _ = _one_step(full_setup, inp, 10.0, "plain", scale=1.0)
    _record("full 25x72x128 sigma=10 plain: 13 residuals", errs)
    _record("full 25x72x128 sigma=10 plain: noise_pred", e_pred)
    assert max(errs) < TOL_RES, errs
    assert e_pred < TOL, e_pred
