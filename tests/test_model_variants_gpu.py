"""GPU parity of the remaining rows of SURVEY.md §8a / §8b: the bbox tower (P2c), calls without ControlNet
conditioning / residuals, single-frame calls (the training loop's F = 1 aux pass), from_unet, and the argument errors
the mirrors raise like the reference."""
import math

import pytest
import torch

from parity_util import make_small_inputs, oracle_pair, rel_l2, small_cfg

pytestmark = pytest.mark.gpu
TOL = 1e-2
TOL_RES = 2e-2


def _x(inp, sigma):
    x = torch.cat([inp["latents"]] * 2) / (sigma ** 2 + 1) ** 0.5
    return torch.cat([x, inp["image_latents"]], dim=2)


def test_bbox_tower_parity(cuda_dev):
    """controlnet_sdv_bbox.py:109-138,551: second conditioning tower, projected with the SHARED conv_out."""
    from posetraj_b200.models import ControlNetSDVModel
    cfg = small_cfg()
    _, o_cnet = oracle_pair(cfg, seed=11, bbox=True)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev, bbox=True)
    inp = make_small_inputs(cfg)
    x, t = _x(inp, 3.0), torch.tensor(0.25 * math.log(3.0))
    bbox_img = inp["controlnet_condition"].flip(-1).contiguous()
    with torch.no_grad():
        o_down, o_mid = o_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"],
                               controlnet_cond=inp["controlnet_condition"], controlnet_bbox=bbox_img)
        o_emb = o_cnet.controlnet_cond_embedding(inp["controlnet_condition"], None, bbox_img)
        o_emb0 = o_cnet.controlnet_cond_embedding(inp["controlnet_condition"], None, None)
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    down, mid = cnet(x.to(cuda_dev), t.to(cuda_dev), d["image_embeddings"], d["added_time_ids"],
                     controlnet_cond=d["controlnet_condition"], controlnet_bbox=bbox_img.to(cuda_dev), return_dict=False)
    torch.cuda.synchronize()
    errs = [rel_l2(a, b) for a, b in zip(down + [mid], o_down + [o_mid])]
    assert max(errs) < TOL_RES, errs
    # the conditioning embedding itself (what is added to conv_in's output, controlnet_sdv.py:599) with both towers
    plan = next(iter(cnet._plans.values()))
    n, (h, w) = plan.n, plan.level_hw[0]
    emb = plan.cond_emb.view(n, h, w, -1).permute(0, 3, 1, 2)
    assert rel_l2(emb, o_emb) < TOL
    assert rel_l2(o_emb, o_emb0) > 0.1                   # the second tower really contributes


def test_no_conditioning_and_no_residuals(cuda_dev):
    """controlnet_cond=None (controlnet_sdv.py:596 guard) and a plain UNet call without residuals."""
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=12)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev)
    inp = make_small_inputs(cfg)
    x, t = _x(inp, 30.0), torch.tensor(0.25 * math.log(30.0))
    with torch.no_grad():
        o_down, o_mid = o_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=None)
        o_plain = o_unet(x, t, inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    out = cnet(x.to(cuda_dev), t.to(cuda_dev), d["image_embeddings"], d["added_time_ids"], controlnet_cond=None)
    assert max(rel_l2(a, b) for a, b in zip(list(out.down_block_res_samples) + [out.mid_block_res_sample], o_down + [o_mid])) < TOL_RES
    plain = unet(x.to(cuda_dev), t.to(cuda_dev), d["image_embeddings"], added_time_ids=d["added_time_ids"]).sample
    assert rel_l2(plain, o_plain) < TOL
    # float timestep (the pipeline passes a 0-dim tensor, training code a python float) gives the same result
    plain2 = unet(x.to(cuda_dev), float(t), d["image_embeddings"], added_time_ids=d["added_time_ids"], return_dict=False)[0]
    assert torch.equal(plain, plain2)


def test_single_frame_call(cuda_dev):
    """F = 1 (the training loop's "spatial" aux pass, scripts/train_svd_traj_VIPSeg_14.py:1396-1404)."""
    from posetraj_b200.models import UNetSpatioTemporalConditionControlNetModel
    cfg = small_cfg(num_frames=1)
    o_unet, _ = oracle_pair(cfg, seed=13)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    inp = make_small_inputs(cfg)
    x, t = _x(inp, 1.0), torch.tensor(0.0)
    with torch.no_grad():
        want = o_unet(x, t, inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
    got = unet(x.to(cuda_dev), t.to(cuda_dev), inp["image_embeddings"].to(cuda_dev),
               added_time_ids=inp["added_time_ids"].to(cuda_dev)).sample
    assert got.shape == want.shape and rel_l2(got, want) < TOL


def test_from_unet_copies_the_encoder_not_add_embedding(cuda_dev):
    """controlnet_sdv.py:653-709: conv_in, time_embedding, down_blocks, mid_block are copied; add_embedding is not."""
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    cfg = small_cfg()
    unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, cuda_dev, seed=4)
    cnet = ControlNetSDVModel.from_unet(unet)
    usd, csd = unet.state_dict(), cnet.state_dict()
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.1.resnets.0.spatial_res_block.conv1.weight",
              "mid_block.attentions.0.proj_in.weight"):
        assert torch.equal(usd[k], csd[k]), k
    assert not torch.equal(usd["add_embedding.linear_1.weight"], csd["add_embedding.linear_1.weight"])
    assert float(csd["controlnet_down_blocks.3.weight"].abs().max()) == 0.0      # zero_module
    assert float(csd["controlnet_cond_embedding.conv_out.weight"].abs().max()) == 0.0
    # faithful init: all residuals are exactly zero (SURVEY.md §4 invariant i)
    inp = make_small_inputs(cfg)
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    down, mid = cnet(_x(inp, 5.0).to(cuda_dev), torch.tensor(0.4, device=cuda_dev), d["image_embeddings"], d["added_time_ids"],
                     controlnet_cond=d["controlnet_condition"], return_dict=False)
    assert all(float(r.float().abs().max()) == 0.0 for r in down) and float(mid.float().abs().max()) == 0.0


def test_argument_errors(cuda_dev):
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    cfg = small_cfg()
    unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, cuda_dev, seed=5)
    cnet = ControlNetSDVModel.from_random(cfg, cuda_dev, seed=5)
    inp = make_small_inputs(cfg)
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    x, t = _x(inp, 5.0).to(cuda_dev), torch.tensor(0.4, device=cuda_dev)
    with pytest.raises(ValueError):     # cross-attention on this path is one token per row
        unet(x, t, d["image_embeddings"].repeat(1, 2, 1), added_time_ids=d["added_time_ids"])
    with pytest.raises(ValueError):     # camera_cond on a model without cc_projection
        cnet(x, t, d["image_embeddings"], d["added_time_ids"], controlnet_cond=d["controlnet_condition"], camera_cond=d["camera_cond"])
    with pytest.raises(ValueError):     # wrong number of residuals
        unet(x, t, d["image_embeddings"], down_block_additional_residuals=[torch.zeros(1, device=cuda_dev)] * 3,
             mid_block_additional_residual=None, added_time_ids=d["added_time_ids"])
    with pytest.raises(RuntimeError):   # CPU tensors: no fallback
        unet(x.cpu(), t.cpu(), inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
    with pytest.raises(ValueError):     # latent size must survive three stride-2 levels
        unet(x[..., :12, :20], t, d["image_embeddings"], added_time_ids=d["added_time_ids"])
