"""ORACLE (test infrastructure, never the product path) — CPU restatement of ONE training step of the reference
(SURVEY.md §8f row 4 / BASELINE configs[3]), written before any backward kernel exists so that the kernels have
something to be checked against.  Everything cites /root/reference/scripts/train_svd_traj_VIPSeg_14_cam_concat.py:

  stratified_uniform / rand_cosine_interpolated   :289-336   (pinned: tests/golden/train_sigma_golden.json holds outputs
                                                              of the reference's OWN two functions, cut out by AST)
  the step                                        :1320-1475
      noise, sigma ~ rand_cosine_interpolated(image_d 64, noise_d 32..64, sigma_data .5, 0.002..700)     :1323-1328
      conditional latents (z0 + 0.02 eps0) / scaling_factor                                                :1336-1339
      noisy = z + eps * sigma ; t = 0.25 ln sigma ; input = noisy / sqrt(sigma^2 + 1)                      :1342-1346
      added_time_ids = [fps 6, noise_aug 0.02, motion]   (NOT the inference order [6, 128, 0.02])         :1352-1361,1223-1255
      conditioning dropout masks                                                                           :1365-1385
      ControlNet -> UNet -> denoised = v c_out + c_skip noisy ; EDM-weighted MSE                           :1404-1436
      "spatial" auxiliary pass: UNet on ONE random frame (F = 1) with residuals sliced `sample[ran_idx]` on the
      flattened B*F axis (only meaningful for batch size 1 — replicated literally), weight 0.5            :1438-1462
      backward through the ControlNet only (the UNet is frozen)                                            :1101,1470

PARITY UNPINNED for the network part (same status as oracle/models.py); the sigma sampler is pinned.
Only tests/ may import this.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

MIN_VALUE, MAX_VALUE, IMAGE_D, NOISE_D_LOW, NOISE_D_HIGH, SIGMA_DATA = 0.002, 700, 64, 32, 64, 0.5   # :338-343
TRAIN_NOISE_AUG = 0.02                                                                               # :1335


def stratified_uniform(shape, group=0, groups=1, dtype=None, device=None, generator=None):
    """:289-299 (k-diffusion): one uniform draw per stratum of [0, 1)."""
    if groups <= 0:
        raise ValueError(f"groups must be positive, got {groups}")
    if group < 0 or group >= groups:
        raise ValueError(f"group must be in [0, {groups})")
    n = shape[-1] * groups
    offsets = torch.arange(group, n, groups, dtype=dtype, device=device)
    u = torch.rand(shape, dtype=dtype, device=device, generator=generator)
    return (offsets + u) / n


def logsnr_to_sigma(u: torch.Tensor, image_d=IMAGE_D, noise_d_low=NOISE_D_LOW, noise_d_high=NOISE_D_HIGH,
                    sigma_data=SIGMA_DATA, min_value=MIN_VALUE, max_value=MAX_VALUE) -> torch.Tensor:
    """The deterministic part of `rand_cosine_interpolated` (:302-336): u in [0,1) -> sigma."""
    logsnr_min = -2 * math.log(min_value / sigma_data)
    logsnr_max = -2 * math.log(max_value / sigma_data)

    def cosine(t, lo, hi):
        t_min = math.atan(math.exp(-0.5 * hi))
        t_max = math.atan(math.exp(-0.5 * lo))
        return -2 * torch.log(torch.tan(t_min + t * (t_max - t_min)))

    def shifted(t, noise_d):
        shift = 2 * math.log(noise_d / image_d)
        return cosine(t, logsnr_min - shift, logsnr_max - shift) + shift

    logsnr = torch.lerp(shifted(u, noise_d_low), shifted(u, noise_d_high), u)
    return torch.exp(-logsnr / 2) * sigma_data


def rand_cosine_interpolated(shape, generator=None, dtype=torch.float32, **kw) -> torch.Tensor:
    return logsnr_to_sigma(stratified_uniform(shape, dtype=dtype, generator=generator), **kw)


def add_time_ids(fps, motion_bucket_ids: torch.Tensor, noise_aug_strength, batch_size: int) -> torch.Tensor:
    """:1223-1255 — [fps, noise_aug_strength, motion] per sample (note the order)."""
    m = motion_bucket_ids.reshape(-1, 1).to(torch.float32)
    if m.shape[0] != batch_size:
        raise ValueError("The length of motion_bucket_ids must match the batch_size.")
    base = torch.tensor([fps, noise_aug_strength], dtype=torch.float32, device=m.device).repeat(batch_size, 1)
    return torch.cat([base, m], dim=1)


def dropout_masks(random_p: torch.Tensor, p: float):
    """:1365-1385 — (prompt_mask [b,1,1] bool, image_mask [b,1,1,1] float) from one uniform draw per sample."""
    bsz = random_p.shape[0]
    prompt_mask = (random_p < 2 * p).reshape(bsz, 1, 1)
    image_mask = 1 - ((random_p >= p).float() * (random_p < 3 * p).float())
    return prompt_mask, image_mask.reshape(bsz, 1, 1, 1)


def training_step(unet, controlnet, *, latents: torch.Tensor, noise: torch.Tensor, sigmas: torch.Tensor,
                  image_embeddings: torch.Tensor, trajectories: torch.Tensor, motion_values: torch.Tensor,
                  camera_cond: Optional[torch.Tensor] = None, scaling_factor: float = 0.18215,
                  random_p: Optional[torch.Tensor] = None, conditioning_dropout_prob: Optional[float] = None,
                  ran_idx: Optional[int] = 0, use_spatial: bool = True) -> Dict[str, torch.Tensor]:
    """One forward of the training step on already-encoded inputs (VAE / CLIP are frozen encoders outside it).

    latents [b, F, 4, h, w] (already multiplied by scaling_factor, :504-512), noise like latents, sigmas [b],
    image_embeddings [b, 1, D], trajectories [b, F, 3, 8h, 8w], motion_values [b].  Returns the losses; call
    `.backward()` on `loss` for the ControlNet gradients (the UNet's parameters are frozen by the caller)."""
    b, F = latents.shape[:2]
    s = sigmas.reshape(b, 1, 1, 1, 1)
    cond = (latents + noise * TRAIN_NOISE_AUG)[:, 0] / scaling_factor
    noisy = latents + noise * s
    timesteps = torch.tensor([0.25 * float(x.log()) for x in sigmas], device=sigmas.device)
    inp = noisy / ((s ** 2 + 1) ** 0.5)
    ehs = image_embeddings
    ids = add_time_ids(6, motion_values, TRAIN_NOISE_AUG, b)
    if conditioning_dropout_prob is not None:
        prompt_mask, image_mask = dropout_masks(random_p, conditioning_dropout_prob)
        ehs = torch.where(prompt_mask, torch.zeros_like(ehs), ehs)
        cond = image_mask * cond
    inp = torch.cat([inp, cond.unsqueeze(1).repeat(1, F, 1, 1, 1)], dim=2)
    kw = {} if camera_cond is None else {"camera_cond": camera_cond}
    down, mid = controlnet(inp, timesteps, ehs, added_time_ids=ids, controlnet_cond=trajectories, return_dict=False, **kw)
    pred = unet(inp, timesteps, ehs, added_time_ids=ids, down_block_additional_residuals=list(down),
                mid_block_additional_residual=mid)
    pred = pred[0] if isinstance(pred, tuple) else getattr(pred, "sample", pred)
    c_out = -s / ((s ** 2 + 1) ** 0.5)
    c_skip = 1 / (s ** 2 + 1)
    weighing = (1 + s ** 2) * (s ** -2.0)
    denoised = pred * c_out + c_skip * noisy
    loss_main = torch.mean((weighing * (denoised - latents) ** 2).reshape(b, -1), dim=1).mean()
    out = {"loss_main": loss_main, "model_pred": pred}
    loss = loss_main
    if use_spatial:
        # second UNet pass on one frame; the residuals are indexed on the FLATTENED (b*F) axis, as the reference does
        sp_in = inp[:, ran_idx].unsqueeze(1)
        sp = unet(sp_in, timesteps, ehs, added_time_ids=ids,
                  down_block_additional_residuals=[r[ran_idx].unsqueeze(0) for r in down],
                  mid_block_additional_residual=mid[ran_idx].unsqueeze(0))
        sp = sp[0] if isinstance(sp, tuple) else getattr(sp, "sample", sp)
        s4 = s[:, 0]
        den_sp = sp[:, 0] * c_out[:, 0] + c_skip[:, 0] * noisy[:, ran_idx]
        loss_sp = torch.mean((weighing[:, 0] * (den_sp - latents[:, ran_idx]) ** 2).reshape(b, -1), dim=1).mean()
        out["loss_spatial"] = loss_sp
        loss = loss + 0.5 * loss_sp
        del s4
    out["loss"] = loss
    return out
