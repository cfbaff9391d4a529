"""Shared helpers of the parity tests: a small SVD-shaped config, oracle <-> product weight hand-over."""
import torch

SMALL = dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=256,
             num_frames=3)


def small_cfg(**over):
    from posetraj_b200.config import SVDConfig
    kw = dict(SMALL)
    kw.update(over)
    return SVDConfig(**kw)


def oracle_pair(cfg, seed=0, cam=False, bbox=False, randomize_zero_convs=True):
    """Oracle models whose weights are rounded to bf16 values (so both sides hold IDENTICAL weights)."""
    from oracle.models import build_models
    unet, cnet = build_models(seed=seed, cam=cam, bbox=bbox, randomize_zero_convs=randomize_zero_convs,
                              in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                              block_out_channels=cfg.block_out_channels,
                              addition_time_embed_dim=cfg.addition_time_embed_dim,
                              projection_class_embeddings_input_dim=cfg.projection_class_embeddings_input_dim,
                              layers_per_block=cfg.layers_per_block, cross_attention_dim=cfg.cross_attention_dim,
                              num_attention_heads=cfg.num_attention_heads, num_frames=cfg.num_frames)
    with torch.no_grad():
        for m in (unet, cnet):
            for name, p in m.named_parameters():
                if p.dim() > 1:  # matrices/kernels live in bf16 on the device; vectors stay fp32
                    p.copy_(p.to(torch.bfloat16).float())
    return unet, cnet


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def make_small_inputs(cfg, h=16, w=24, seed=1234):
    from oracle.pipeline import make_inputs
    inp = make_inputs(num_frames=cfg.num_frames, h=h, w=w, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    inp["image_embeddings"] = torch.cat([torch.zeros(1, 1, cfg.cross_attention_dim),
                                         torch.randn(1, 1, cfg.cross_attention_dim, generator=g)])
    H, W = h * 8, w * 8
    cond = torch.full((cfg.num_frames, 3, H, W), -1.0)
    # a synthetic "trajectory drawing": a few bright blobs moving across frames
    for f in range(cfg.num_frames):
        y0, x0 = 20 + 9 * f, 30 + 14 * f
        cond[f, 0, y0:y0 + 7, x0:x0 + 40] = 1.0
        cond[f, 1, y0 + 3:y0 + 10, x0 + 36:x0 + 43] = 1.0
    inp["controlnet_condition"] = torch.cat([cond[None]] * 2)
    inp["camera_cond"] = torch.cat([torch.randn(1, cfg.num_frames, 12, generator=g) * 0.1] * 2)
    return inp


def make_small_bbox_maps(cfg, inp, seed=4321):
    """Sparse +-1 maps for the bbox tower (controlnet_sdv_bbox.py:109-138), same for both rows of the CFG pair."""
    g = torch.Generator().manual_seed(seed)
    hw = inp["controlnet_condition"].shape[-2:]
    return torch.cat([(torch.rand(1, cfg.num_frames, 3, *hw, generator=g) > 0.97).float() * 2 - 1] * 2)
