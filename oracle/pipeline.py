"""ORACLE (test infrastructure) — the denoise loop of
/root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:526-583 (cam: ..._cam.py:532-590),
latent-to-latent (VAE / CLIP excluded).
"""
from __future__ import annotations

import torch

from .scheduler import EulerKarrasOracle


def make_inputs(num_frames=14, h=40, w=72, seed=1234, cond_hw=None, dtype=torch.float32):
    """Seeded synthetic inputs in the order SURVEY.md §8(d) lists them."""
    g = torch.Generator().manual_seed(seed)
    init_sigma = (700.0 ** 2 + 1) ** 0.5
    latents = torch.randn(1, num_frames, 4, h, w, generator=g) * init_sigma
    img = torch.randn(1, 4, h, w, generator=g)
    image_latents = torch.cat([torch.zeros_like(img), img])[:, None].repeat(1, num_frames, 1, 1, 1)
    emb = torch.randn(1, 1, 1024, generator=g)
    image_embeddings = torch.cat([torch.zeros_like(emb), emb])
    H, W = cond_hw if cond_hw is not None else (h * 8, w * 8)
    added_time_ids = torch.tensor([[6.0, 128.0, 0.02]] * 2)
    guidance = torch.linspace(1.0, 3.0, num_frames)
    return dict(latents=latents.to(dtype), image_latents=image_latents.to(dtype),
                image_embeddings=image_embeddings.to(dtype), added_time_ids=added_time_ids.to(dtype),
                guidance=guidance.to(dtype), cond_hw=(H, W))


@torch.no_grad()
def denoise_step(unet, controlnet, sched, latents, i, t, image_latents, image_embeddings, controlnet_condition,
                 added_time_ids, guidance, cond_scale=1.0, camera_cond=None, return_pred=False):
    """One iteration of the loop (:530-572)."""
    x = torch.cat([latents] * 2)
    x = sched.scale_model_input(x, t)
    x = torch.cat([x, image_latents], dim=2)
    kw = {} if camera_cond is None else {"camera_cond": camera_cond}
    down, mid = controlnet(x, t, encoder_hidden_states=image_embeddings, controlnet_cond=controlnet_condition,
                           added_time_ids=added_time_ids, conditioning_scale=cond_scale, guess_mode=False,
                           return_dict=False, **kw)
    pred = unet(x, t, encoder_hidden_states=image_embeddings, down_block_additional_residuals=down,
                mid_block_additional_residual=mid, added_time_ids=added_time_ids, return_dict=False)
    u, c = pred.chunk(2)
    g = guidance.view(1, -1, 1, 1, 1).to(pred.dtype)
    noise_pred = u + g * (c - u)
    new_latents = sched.step(noise_pred, t, latents)
    if return_pred:
        return new_latents, pred
    return new_latents


@torch.no_grad()
def denoise(unet, controlnet, latents, image_latents, image_embeddings, controlnet_condition, added_time_ids,
            guidance, num_inference_steps=25, cond_scale=1.0, camera_cond=None, max_steps=None):
    sched = EulerKarrasOracle()
    sched.set_timesteps(num_inference_steps)
    for i, t in enumerate(sched.timesteps):
        if max_steps is not None and i >= max_steps:
            break
        latents = denoise_step(unet, controlnet, sched, latents, i, t, image_latents, image_embeddings,
                               controlnet_condition, added_time_ids, guidance, cond_scale, camera_cond)
    return latents
