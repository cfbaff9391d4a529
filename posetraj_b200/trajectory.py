"""Trajectory conditioning maps on the GPU.

Reference: the CPU/OpenCV loop of /root/reference/scripts/run_inference_vipseg_json_repro.py:426-449 (identical drawing
in /root/reference/utils/dataset.py:741-766 and scripts/train_svd_traj_VIPSeg_14.py:204-215): tracks are rescaled to the
target size with `int()` truncation, each frame transition k gets a black canvas with a red 3-px line p_k -> p_k+1 and
a green radius-3 disc at p_k+1 per track (painter's order), a black image is appended, and the pipeline turns the RGB
images into a [F, 3, H, W] tensor in [-1, 1] (pipeline_stable_video_diffusion_controlnet.py:500).

`rasterize_tracks` does all of that in two kernels of libposetraj_b200.so (csrc/raster.cu), bit-exact against OpenCV.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import torch

from . import _lib


def rescale_tracks(trajectory_json: Dict[str, Sequence[Sequence[float]]], size: Sequence[int],
                   original_size: Sequence[int], style: str = "inference") -> List[List[List[int]]]:
    """`size` = [height, width] of the target, `original_size` = (height, width, ...) of the source frames.

    style "inference": int(x * (W / W0)), int(y * (H / H0))   — run_inference_vipseg_json_repro.py:431
    style "dataset"  : int(x / W0 * W),   int(y / H0 * H)     — utils/dataset.py:750
    (Python floats are IEEE doubles: the two orders can differ by one pixel, so both are kept.)"""
    out = []
    for index in trajectory_json:
        pts = trajectory_json[index]
        if style == "inference":
            out.append([[int(i[0] * (size[1] / original_size[1])), int(i[1] * (size[0] / original_size[0]))] for i in pts])
        elif style == "dataset":
            out.append([[int(i[0] / original_size[1] * size[1]), int(i[1] / original_size[0] * size[0])] for i in pts])
        else:
            raise ValueError("style must be 'inference' or 'dataset'")
    return out


def rasterize_tracks(tracks, num_frames: int, height: int, width: int, device=None, output: str = "f32",
                     start: int = 0, style: str = "inference") -> torch.Tensor:
    """tracks: K lists (or an int tensor [K, >= start+num_frames, 2]) of (x, y) pixel coordinates.

    Returns the `num_frames` conditioning maps: output "f32" -> [F, 3, H, W] float32 in [-1, 1] (what the pipeline
    feeds the ControlNet), "u8" -> [F, H, W, 3] uint8 RGB (the PIL images of the reference).  Frame k draws the motion
    k -> k+1; the last frame is black.  style "dataset" reproduces utils/dataset.py:741-766, whose BGR->RGB conversion
    sits inside the track loop (the channels swap once per track; that variant also has no black padding frame in the
    reference, the caller drops it)."""
    if style not in ("inference", "dataset"):
        raise ValueError("style must be 'inference' or 'dataset'")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("posetraj_b200: the rasteriser runs on CUDA only (no CPU fallback)")
    if output not in ("f32", "u8"):
        raise ValueError("output must be 'f32' or 'u8'")
    t = torch.as_tensor(tracks, dtype=torch.int32) if not torch.is_tensor(tracks) else tracks.to(torch.int32)
    if t.numel() == 0:
        t = torch.zeros(0, num_frames, 2, dtype=torch.int32)
    if t.dim() != 3 or t.shape[2] != 2 or t.shape[1] < start + num_frames:
        raise ValueError(f"tracks must be [K, >= {start + num_frames}, 2], got {tuple(t.shape)}")
    t = t[:, start:start + num_frames].contiguous().to(device)
    K = t.shape[0]
    lib = _lib.lib()
    ws = lib.pt_rasterize_workspace_bytes(num_frames, height, width)
    if ws <= 0:
        raise ValueError("bad raster shape")
    order = torch.empty(ws // 4, device=device, dtype=torch.int32)
    if output == "f32":
        out = torch.empty(num_frames, 3, height, width, device=device, dtype=torch.float32)
    else:
        out = torch.empty(num_frames, height, width, 3, device=device, dtype=torch.uint8)
    a = _lib.PtRasterArgs()
    a.tracks = t.data_ptr() if K > 0 else None
    a.K, a.F, a.H, a.W = K, num_frames, height, width
    a.order = order.data_ptr()
    a.out_f32 = out.data_ptr() if output == "f32" else None
    a.out_u8 = out.data_ptr() if output == "u8" else None
    a.swap_per_track = 1 if style == "dataset" else 0
    with torch.cuda.device(device):
        _lib.check(lib.pt_rasterize_tracks(C.addressof(a), torch.cuda.current_stream().cuda_stream), "pt_rasterize_tracks")
    return out
