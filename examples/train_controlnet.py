"""The core of the reference's training loop (scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1320-1475) on posetraj_b200:
sigma sampling, one ControlNetTrainer.step per batch (both forwards, EDM loss, spatial pass, backward, all-reduce, AdamW),
checkpoint in the diffusers layout.  Data here is synthetic (latents / embeddings / trajectory maps of the right shapes): the
reference's dataset, VAE and CLIP encoding stay the caller's (`posetraj_b200.AutoencoderKLTemporalDecoder` /
`CLIPVisionModelWithProjection` provide the encoders).

    python examples/train_controlnet.py --steps 5 --small
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_controlnet.py --steps 100     # data-parallel, 2 videos per GPU
"""
import argparse
import math
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from posetraj_b200 import ControlNetSDVModel, ControlNetTrainer, SVDConfig, UNetSpatioTemporalConditionControlNetModel  # noqa: E402


def rand_cosine_interpolated(n, device, image_d=64, noise_d_low=32, noise_d_high=64, sigma_data=0.5, min_value=0.002, max_value=700):
    """The reference's sigma sampler (train...cam_concat.py:289-336: stratified uniform -> interpolated cosine log-SNR)."""
    u = (torch.arange(n, device=device, dtype=torch.float32) + torch.rand(n, device=device)) / n
    lo, hi = -2 * math.log(min_value / sigma_data), -2 * math.log(max_value / sigma_data)

    def shifted(t, noise_d):
        shift = 2 * math.log(noise_d / image_d)
        t_min, t_max = math.atan(math.exp(-0.5 * (hi - shift))), math.atan(math.exp(-0.5 * (lo - shift)))
        return -2 * torch.log(torch.tan(t_min + t * (t_max - t_min))) + shift

    logsnr = torch.lerp(shifted(u, noise_d_low), shifted(u, noise_d_high), u)
    return torch.exp(-logsnr / 2) * sigma_data


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=2, help="videos per GPU")
    ap.add_argument("--frames", type=int, default=14)
    ap.add_argument("--height", type=int, default=40, help="latent height (pixels / 8)")
    ap.add_argument("--width", type=int, default=72)
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--bbox", action="store_true", help="controlnet_sdv_bbox: second conditioning tower")
    ap.add_argument("--small", action="store_true", help="a SMALL config (smoke runs)")
    ap.add_argument("--out", default=None, help="directory for the trained ControlNet (diffusers layout)")
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    cfg = SVDConfig(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=256,
                    num_frames=args.frames) if args.small else SVDConfig(num_frames=args.frames)
    unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)            # frozen (:984-987)
    controlnet = ControlNetSDVModel.from_random(cfg, dev, seed=1, bbox=args.bbox, faithful_zero_init=False)
    trainer = ControlNetTrainer(unet, controlnet, batch=args.batch, frames=args.frames, height=args.height, width=args.width,
                                lr=args.lr)                                                   # AdamW as :1113-1123
    trainer.use_cuda_graph = True
    b, F, h, w = args.batch, args.frames, args.height, args.width
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    random.seed(rank)
    losses, t0 = [], None
    for step in range(args.steps):
        # --- what the reference's dataloader + frozen encoders deliver (:1311-1340), here synthetic -------------
        latents = torch.randn(b, F, 4, h, w, device=dev, generator=g) * 0.9                   # VAE latents x scaling_factor
        image_embeddings = torch.randn(b, 1, cfg.cross_attention_dim, device=dev, generator=g)
        maps = (torch.rand(b, F, 3, 8 * h, 8 * w, device=dev, generator=g) > 0.97).float() * 2 - 1
        bbox = (torch.rand(b, F, 3, 8 * h, 8 * w, device=dev, generator=g) > 0.98).float() * 2 - 1 if args.bbox else None
        # --- the step (:1320-1475) -------------------------------------------------------------------------------
        noise = torch.randn(latents.shape, device=dev, generator=g)
        sigmas = rand_cosine_interpolated(b, dev)
        loss = trainer.step(latents=latents, noise=noise, sigmas=sigmas, image_embeddings=image_embeddings, trajectories=maps,
                            motion_values=torch.full((b,), 127.0, device=dev), controlnet_bbox=bbox, ran_idx=random.randint(0, F - 1))
        losses.append(loss.clone())
        if step == min(2, args.steps - 1):
            torch.cuda.synchronize()
            t0, s0 = time.time(), step
    torch.cuda.synchronize()
    if rank == 0:
        ls = [float(x) for x in losses]
        if args.steps - 1 > s0:
            ms = (time.time() - t0) / (args.steps - 1 - s0) * 1e3
            print(f"{ms:.1f} ms/step ({b * world / ms * 1e3:.1f} videos/s on {world} GPU(s)); loss {ls[0]:.4f} -> {ls[-1]:.4f}")
        else:
            print(f"loss {ls[0]:.4f} -> {ls[-1]:.4f}")
        if args.out:
            print("saved", trainer.save_pretrained(os.path.join(args.out, "controlnet")))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return losses


if __name__ == "__main__":
    main()
