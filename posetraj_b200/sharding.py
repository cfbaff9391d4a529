"""Multi-GPU partitioning of the denoise path.  One process per GPU (torchrun); NCCL only where the
path has a real exchange step.

  * video batch  — videos are independent: `video_shard` hands each rank a contiguous slice, no data-path collective;
                   `gather_latents` collects the final latents on every rank (one all-gather per call).
  * CFG branch   — one video on 2 GPUs: rank r computes row r of the CFG pair through ControlNet + UNet
                   (`StableVideoDiffusionPipelineControlNet.enable_cfg_split`), the two 161 KB predictions are
                   all-gathered each step, both ranks redo the CFG + Euler update.  Legal because activations never
                   cross batch rows; the temporal cross-attention context DOES depend on every row's image embedding
                   (reference quirk), so each rank keeps both embeddings and
                   `temporal_context_rotation` says how its lookup table is rotated.
  * frames       — one video on N GPUs: posetraj_b200/frame_sharding.py (frame-sharded spatial layers, pixel-sharded
                   temporal layers, all-to-all around every temporal sub-block); `frame_shards` fixes the ragged
                   partition (F = 25 is not divisible by 2/4/8).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def video_shard(num_videos: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced slice of `num_videos` for `rank` (the first `num_videos % world_size` ranks get one more)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(num_videos, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def frame_shards(num_frames: int, world_size: int) -> List[Tuple[int, int]]:
    """(start, count) per rank for frame-sharded spatial layers; ragged when world_size does not divide F."""
    base, extra = divmod(num_frames, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def temporal_context_rotation(row: int, hw: int, ctx_batch: int = 2) -> int:
    """A shard holding only batch row `row` sees local row 0, while the reference indexes the temporal cross-attention
    context of hidden row (b, s) by (b*HW + s) mod B: the shard's table must be rotated by (row*HW) mod B."""
    return (row * hw) % ctx_batch


def cfg_split_ranks(world_size: int) -> List[Tuple[int, int]]:
    """Pairs of ranks that share one video under CFG-branch sharding: [(uncond_rank, cond_rank), ...]."""
    if world_size % 2:
        raise ValueError("CFG-branch sharding needs an even number of ranks")
    return [(2 * i, 2 * i + 1) for i in range(world_size // 2)]


def gather_latents(latents: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of per-rank final latents [n_local, F, C, h, w] (equal n_local on every rank) -> [world*n_local, ...].
    Works on NCCL (GPU) and gloo (CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [torch.empty_like(latents) for _ in range(world)]
    dist.all_gather(out, latents.contiguous(), group=group)
    return torch.cat(out, dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Device-timed numbers are reported as the max over ranks (never wall clock)."""
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
