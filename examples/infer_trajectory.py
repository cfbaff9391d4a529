"""The core of the reference's inference script (scripts/run_inference_vipseg_json_repro.py:335-341, 420-451) on
posetraj_b200: load the models, read a CoTracker trajectory JSON, draw the trajectory maps (on the GPU, cv2-exact), run
the pipeline, write the frames.

    python examples/infer_trajectory.py --json tests/golden/traj_9_E0zfiF4DCt8.json --out /tmp/frames.npz
    python examples/infer_trajectory.py --svd /ckpt/stable-video-diffusion-img2vid --controlnet /ckpt/posetraj --image in.png ...

Without --svd / --controlnet the networks are random-init SVD-shaped (no checkpoint ships with this repository): the
script then demonstrates the call sequence and the timing, not a meaningful video."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from posetraj_b200 import (AutoencoderKLTemporalDecoder, CLIPVisionModelWithProjection, ControlNetSDVModel,  # noqa: E402
                           StableVideoDiffusionPipelineControlNet, SVDConfig, UNetSpatioTemporalConditionControlNetModel)
from posetraj_b200.trajectory import rasterize_tracks, rescale_tracks  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--svd", default=None, help="diffusers directory of stable-video-diffusion-img2vid (unet/, vae/, image_encoder/)")
    ap.add_argument("--controlnet", default=None, help="directory holding controlnet/ (diffusers layout)")
    ap.add_argument("--json", required=True, help="CoTracker trajectories: {track id: [[x, y] per frame]}")
    ap.add_argument("--image", default=None, help="first frame (any PIL-readable file); default: a grey test card")
    ap.add_argument("--original-size", type=int, nargs=2, default=[720, 1280], help="(height, width) the trajectories refer to")
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=576)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--tracks", type=int, default=0, help="use only the first N tracks (0: all)")
    ap.add_argument("--small", action="store_true", help="random-init models of a SMALL config (smoke runs)")
    ap.add_argument("--out", default="frames.npz")
    args = ap.parse_args(argv)
    dev = torch.device("cuda", 0)

    if args.svd:
        unet = UNetSpatioTemporalConditionControlNetModel.from_pretrained(args.svd, subfolder="unet", variant="fp16", device=dev)
        controlnet = ControlNetSDVModel.from_pretrained(args.controlnet, subfolder="controlnet", device=dev)
        pipe = StableVideoDiffusionPipelineControlNet.from_pretrained(args.svd, controlnet=controlnet, unet=unet, variant="fp16")
    else:
        from posetraj_b200.clip import CLIPVisionConfig
        from posetraj_b200.vae import VaeConfig
        cfg = SVDConfig(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=256) if args.small \
            else SVDConfig()
        unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
        controlnet = ControlNetSDVModel.from_random(cfg, dev, seed=0, faithful_zero_init=False)
        vcfg = VaeConfig(block_out_channels=(64, 64, 128, 128)) if args.small else VaeConfig()
        ccfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, image_size=56,
                                patch_size=14, projection_dim=cfg.cross_attention_dim) if args.small else CLIPVisionConfig()
        pipe = StableVideoDiffusionPipelineControlNet(vae=AutoencoderKLTemporalDecoder.from_random(vcfg, dev), unet=unet,
                                                      image_encoder=CLIPVisionModelWithProjection.from_random(ccfg, dev),
                                                      controlnet=controlnet)
    pipe.enable_model_cpu_offload()          # no-ops kept so that the reference script runs unchanged (:339-341)

    with open(args.json) as f:
        trajectory_json = json.load(f)
    if "tracks" in trajectory_json and isinstance(trajectory_json["tracks"], dict):   # this repository's fixture wraps the CoTracker dict
        args.original_size = trajectory_json.get("assumed_original_size", args.original_size)
        trajectory_json = trajectory_json["tracks"]
    if args.tracks:
        trajectory_json = {k: trajectory_json[k] for k in list(trajectory_json)[: args.tracks]}
    size = [args.height, args.width]
    tracks = rescale_tracks(trajectory_json, size, args.original_size)                    # :431
    maps = rasterize_tracks(tracks, 14, args.height, args.width, dev)                     # :433-449 (cv2.line / cv2.circle), on the GPU
    if args.image:
        from PIL import Image
        image = Image.open(args.image).convert("RGB").resize((args.width, args.height))
    else:
        image = torch.full((1, 3, args.height, args.width), 0.5)
    torch.cuda.synchronize()
    t0 = time.time()
    frames = pipe(image, maps, decode_chunk_size=8, num_frames=14, motion_bucket_id=10, controlnet_cond_scale=1.0,
                  width=args.width, height=args.height, num_inference_steps=args.steps, output_type="np").frames   # :451
    torch.cuda.synchronize()
    dt = time.time() - t0
    video = np.asarray(frames[0])
    np.savez_compressed(args.out, frames=video, trajectory_maps=rasterize_tracks(tracks, 14, args.height, args.width, dev, output="u8").cpu().numpy())
    print(f"{len(tracks)} tracks -> {video.shape} frames in {dt:.2f} s ({args.steps} steps) -> {args.out}")
    return video


if __name__ == "__main__":
    main()
