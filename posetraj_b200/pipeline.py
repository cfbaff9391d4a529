"""Mirror of `StableVideoDiffusionPipelineControlNet` with the denoise loop on the sm_100a kernels.

Reference: /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py
  __call__ :317-340 (signature kept verbatim), prepare_latents :267-299, added_time_ids override :513-523,
  guidance scale :506-509, denoise loop :526-583; the `_cam` variant
  (/root/reference/pipeline/pipeline_stable_video_diffusion_controlnet_cam.py:321,506-509,549) adds `camera_cond`.

What runs where
  * per step: ControlNet plan -> UNet plan -> fused CFG + Euler + next-input kernel -> device step counter.  The
    per-step kernel sequence is captured into a CUDA graph (same kernels, replayed) unless a step-end callback
    needs host control between steps.
  * once per call: Karras sigma table (host float64, like the reference), ControlNet conditioning embedding,
    the 1-token cross-attention vectors.
  * the VAE on either side of the loop is posetraj_b200.vae.AutoencoderKLTemporalDecoder on
    the same kernel library: `_encode_vae_image` (:174-195) for the conditioning image, `decode_latents` (:225-251)
    + `tensor2vid` (:70-82) for output_type "pt" / "np" / "pil".  `image_latents=` may be passed instead of a VAE.
  * the image-conditioning branch is posetraj_b200.clip: the reference's anti-aliased resize
    (:602-712) and the CLIP ViT-H/14 vision tower on the same library (`_encode_image`, :145-172);
    `image_embeddings=` may be passed instead of an image_encoder.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Union

import torch

from . import ops
from .engine import BF16, F32, NetPlan
from .models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
from .scheduler import EulerDiscreteScheduler


@dataclass
class StableVideoDiffusionPipelineOutput:
    frames: Union[List, torch.Tensor]


class VaeImageProcessor:
    """The two calls the reference pipeline makes on diffusers' VaeImageProcessor (pipeline...controlnet.py:143,
    450, 500, 78): `preprocess` (PIL / numpy / tensor -> [N, 3, height, width] fp32 in [-1, 1]) and `postprocess`
    ([N, 3, H, W] in [-1, 1] -> "pt" | "np" | "pil").  Host-side image formatting, not arithmetic of the path."""

    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = vae_scale_factor

    @staticmethod
    def _pil_to_pt(images, height, width):
        import numpy as np
        import PIL.Image
        out = []
        for im in images:
            if height is not None and width is not None and im.size != (width, height):
                im = im.resize((width, height), resample=PIL.Image.LANCZOS)
            out.append(np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0)
        return torch.from_numpy(np.stack(out, 0)).permute(0, 3, 1, 2)

    def preprocess(self, image, height: Optional[int] = None, width: Optional[int] = None) -> torch.Tensor:
        import numpy as np
        try:
            import PIL.Image
            pil_t = PIL.Image.Image
        except ImportError:  # pragma: no cover
            pil_t = ()
        if isinstance(image, pil_t):
            image = [image]
        if isinstance(image, (list, tuple)) and len(image) and isinstance(image[0], pil_t):
            x = self._pil_to_pt(image, height, width)
        elif isinstance(image, np.ndarray) or (isinstance(image, (list, tuple)) and isinstance(image[0], np.ndarray)):
            arr = np.stack(image, 0) if isinstance(image, (list, tuple)) else image
            if arr.ndim == 3:
                arr = arr[None]
            x = torch.from_numpy(arr.astype(np.float32)).permute(0, 3, 1, 2)
        elif torch.is_tensor(image) or (isinstance(image, (list, tuple)) and torch.is_tensor(image[0])):
            x = torch.stack(list(image), 0) if isinstance(image, (list, tuple)) else image
            if x.dim() == 3:
                x = x[None]
            x = x.to(F32)
            if x.min() < 0:      # already in [-1, 1] (diffusers warns and skips the normalisation)
                if height is not None and tuple(x.shape[-2:]) != (height, width):
                    x = torch.nn.functional.interpolate(x, size=(height, width))
                return x
        else:
            raise ValueError("image must be a PIL image, numpy array, tensor, or a list of those")
        if height is not None and tuple(x.shape[-2:]) != (height, width):
            x = torch.nn.functional.interpolate(x, size=(height, width))
        return 2.0 * x - 1.0

    def postprocess(self, image: torch.Tensor, output_type: str = "pil"):
        if output_type not in ("pt", "np", "pil"):
            raise ValueError(f"output_type must be one of 'latent', 'pt', 'np', 'pil', got {output_type!r}")
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return image
        arr = image.cpu().permute(0, 2, 3, 1).float().numpy()
        if output_type == "np":
            return arr
        import PIL.Image
        return [PIL.Image.fromarray(a) for a in (arr * 255).round().astype("uint8")]


def tensor2vid(video: torch.Tensor, processor: VaeImageProcessor, output_type: str = "np"):
    """pipeline...controlnet.py:70-82: [B, C, F, H, W] -> per video, the post-processed frames."""
    outputs = []
    for b in range(video.shape[0]):
        outputs.append(processor.postprocess(video[b].permute(1, 0, 2, 3), output_type))
    return outputs


def _get_add_time_ids(noise_aug_strength, dtype, batch_size, fps=4, motion_bucket_id=128, unet=None):
    """pipeline...controlnet.py:37-59 (module-level helper used for the hard override at :513-523)."""
    add_time_ids = [fps, motion_bucket_id, noise_aug_strength]
    passed = unet.config.addition_time_embed_dim * len(add_time_ids)
    expected = unet.add_embedding.linear_1.in_features
    if expected != passed:
        raise ValueError(f"Model expects an added time embedding vector of length {expected}, but a vector of "
                         f"{passed} was created. The model has an incorrect config.")
    return torch.tensor([add_time_ids], dtype=dtype)


class DenoiseEngine:
    """One video's denoise loop: both network plans wired to shared buffers + the fused scheduler kernel.

    `rows=(begin, count)` selects which rows of the CFG pair this process computes (CFG-branch
    sharding): (0, 2) is the whole pair on one GPU; (r, 1) with a 2-rank `group` runs one branch per GPU, all-gathers
    the two 161 KB noise predictions every step and lets both ranks redo the (2 MB) CFG + Euler update."""

    def __init__(self, unet: UNetSpatioTemporalConditionControlNetModel, controlnet: ControlNetSDVModel,
                 scheduler: EulerDiscreteScheduler, *, frames: int, h: int, w: int, cond_hw: tuple, device,
                 rows: tuple = (0, 2), group=None):
        self.unet, self.controlnet, self.scheduler = unet, controlnet, scheduler
        self.F, self.h, self.w, self.device = frames, h, w, device
        self.row_begin, self.row_count = rows
        self.group = group
        self.split = self.row_count != 2
        cfg = unet.cfg
        self.latents = torch.zeros(frames, cfg.out_channels, h, w, device=device, dtype=F32)
        self.image_latents = torch.zeros(2, frames, cfg.out_channels, h, w, device=device, dtype=F32)
        self.guidance = torch.ones(frames, device=device, dtype=F32)
        self.step_index = torch.zeros(1, device=device, dtype=torch.int32)
        self.sigmas = torch.zeros(1024, device=device, dtype=F32)
        kw = dict(sigmas=self.sigmas, step_index=self.step_index)
        if self.split:
            kw.update(ctx_batch=2, row_offset=self.row_begin)
        self.cplan: NetPlan = controlnet.plan_for(self.row_count, frames, h, w, cond_hw=cond_hw, **kw)
        # Two-stream step (PT_TWO_STREAM, default on): the UNet's encoder never reads the ControlNet's residuals before they are
        # added to its skips, so it runs CONCURRENTLY with the ControlNet on a second stream inside the captured graph (the
        # kernels of levels 2-3 fill a fraction of the 148 SMs, and every persistent GEMM has a partial last wave); the 12
        # injections `skip_i += m_i * r_i` follow the join (models/unet_spatio_temporal_condition_controlnet.py:451-469).
        self.two_stream = os.environ.get("PT_TWO_STREAM", "1") != "0"
        ukw = dict(kw, defer_injection=True) if self.two_stream else kw
        self.uplan: NetPlan = unet.plan_for(self.row_count, frames, h, w, x_in=self.cplan.x_in,
                                            residual_bufs=self.cplan.res, **ukw)
        # the prediction of the whole pair: the UNet plan's own buffer, or the all-gather target when sharded
        self.pred_all = self.uplan.noise_pred if not self.split else torch.zeros(
            2 * frames * h * w, cfg.out_channels, device=device, dtype=BF16)
        common = dict(latents=self.latents, guidance=self.guidance, sigmas=self.sigmas, step_index=self.step_index,
                      next_in=self.cplan.x_in, image_latents=self.image_latents, next_padded=True,
                      row_begin=self.row_begin, row_count=self.row_count)
        self.prepare_op = ops.CfgEuler(noise_pred=None, mode=1, **common)
        self.update_op = ops.CfgEuler(noise_pred=self.pred_all, mode=0, **common)
        if self.two_stream:
            k = self.uplan.split_index
            self.unet_head, self.unet_rest = self.uplan.step_ops[:k], self.uplan.inject_ops + self.uplan.step_ops[k:]
            self.net_ops = self.cplan.step_ops + self.unet_head + self.unet_rest
            self._side = torch.cuda.Stream(device=device)
        else:
            self.net_ops = self.cplan.step_ops + self.uplan.step_ops
        self.tail_ops = [self.update_op, ops.StepAdvance(self.step_index)]
        self.step_ops = self.net_ops + self.tail_ops
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = len(self.step_ops)

    def load(self, *, latents, image_latents, image_embeddings, added_time_ids, guidance, sigmas,
             controlnet_condition, camera_cond=None, cond_scale: float = 1.0) -> None:
        """Stage one video's inputs (host->device copies happen here) and run the step-invariant prologue."""
        sp = torch.cuda.current_stream().cuda_stream
        rb, rc = self.row_begin, self.row_count
        self.latents.copy_(latents.reshape(self.latents.shape))
        self.image_latents.copy_(image_latents.reshape(self.image_latents.shape))
        self.guidance.copy_(guidance.reshape(-1))
        n = sigmas.numel()
        self.sigmas[:n].copy_(sigmas)
        self.step_index.zero_()
        for plan in (self.cplan, self.uplan):
            plan.ehs.copy_(image_embeddings[:, 0, :])            # every shard keeps ALL rows' embeddings (fact 11)
            plan.time_ids.copy_(added_time_ids[rb:rb + rc].reshape(-1))
            NetPlan.run(plan.embed_ops, sp)
        # once per video: the conditioning embedding is always re-run
        self.controlnet.stage_condition(self.cplan, controlnet_condition[rb:rb + rc],
                                        None if camera_cond is None else camera_cond[rb:rb + rc], None, sp)
        self.cplan.set_conditioning_scale(cond_scale)
        self.prepare_op.launch(sp)
        self._latents0 = self.latents.clone()

    def reset(self) -> None:
        """Re-arm the loop on the inputs staged by the last load(): device-side only (3 tiny launches)."""
        self.latents.copy_(self._latents0)
        self.step_index.zero_()
        self.prepare_op.launch(torch.cuda.current_stream().cuda_stream)

    def _exchange(self) -> None:
        """CFG-branch sharding: all-gather the two branches' predictions (rank == row of the CFG pair)."""
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(self.pred_all, self.uplan.noise_pred, group=self.group)
        else:
            # host-staged exchange for non-NCCL process groups (gloo: two test ranks sharing one GPU)
            mine = self.uplan.noise_pred.cpu()
            parts = [torch.empty_like(mine) for _ in range(2)]
            dist.all_gather(parts, mine, group=self.group)
            self.pred_all.copy_(torch.cat(parts, 0))

    def step(self, use_graph: bool = True) -> None:
        """One denoise step (all kernels of ControlNet + UNet + CFG/Euler)."""
        sp = torch.cuda.current_stream().cuda_stream
        if not self.split:
            if use_graph and self.graph is not None:
                self.graph.replay()
            else:
                NetPlan.run(self.step_ops, sp)
            return
        if use_graph and self.graph is not None:
            self.graph.replay()                      # the two networks (this rank's branch)
        else:
            NetPlan.run(self.net_ops, sp)
        self._exchange()
        NetPlan.run(self.tail_ops, sp)

    def capture(self) -> None:
        """Capture the per-step kernel sequence into a CUDA graph (call after at least one eager step).  When
        sharded only the network part is captured; the NCCL exchange and the 2 tail kernels stay eager."""
        if self.graph is not None:
            return
        saved = self.step_index.clone(), self.latents.clone(), self.cplan.x_in.clone()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                if self.two_stream:
                    self._side.wait_stream(s)                         # fork
                    with torch.cuda.stream(self._side):
                        NetPlan.run(self.unet_head, self._side.cuda_stream)
                    NetPlan.run(self.cplan.step_ops, s.cuda_stream)
                    s.wait_stream(self._side)                         # join
                    NetPlan.run(self.unet_rest + ([] if self.split else self.tail_ops), s.cuda_stream)
                else:
                    NetPlan.run(self.net_ops if self.split else self.step_ops, torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().wait_stream(s)
        # capture does not execute, but keep state exactly as before anyway
        self.step_index.copy_(saved[0]); self.latents.copy_(saved[1]); self.cplan.x_in.copy_(saved[2])
        self.graph = g


class StableVideoDiffusionPipelineControlNet:
    """Same constructor components and `__call__` signature as the reference pipeline."""

    def __init__(self, vae=None, image_encoder=None, unet: UNetSpatioTemporalConditionControlNetModel = None,
                 controlnet: ControlNetSDVModel = None, scheduler: EulerDiscreteScheduler = None,
                 feature_extractor=None):
        self.vae, self.image_encoder, self.unet, self.controlnet = vae, image_encoder, unet, controlnet
        self.scheduler = scheduler or EulerDiscreteScheduler()
        self.feature_extractor = feature_extractor
        self.vae_scale_factor = 8 if vae is None else 2 ** (len(vae.config.block_out_channels) - 1)
        self.image_processor = VaeImageProcessor(vae_scale_factor=self.vae_scale_factor)
        self._engines: Dict[tuple, DenoiseEngine] = {}
        self._guidance_scale = None
        self._num_timesteps = 0
        self.use_cuda_graph = True
        self._cfg_rows = (0, 2)
        self._cfg_group = None
        self._frame_shard = None   # (rank, world, group) when one video is frame-sharded over several GPUs

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, controlnet: ControlNetSDVModel = None,
                        unet: UNetSpatioTemporalConditionControlNetModel = None, vae=None, image_encoder=None, scheduler=None,
                        feature_extractor=None, variant: Optional[str] = None, device=None, torch_dtype=None, **unused):
        """The constructor call of the reference scripts, unchanged
        (scripts/run_inference_vipseg_json_repro.py:338: `from_pretrained(path, controlnet=controlnet, unet=unet)`): a
        diffusers pipeline directory with `unet/`, `vae/`, `image_encoder/`, `scheduler/` sub-folders.  Components passed
        in are used as they are; the others are loaded from their sub-folder when it exists (a missing `vae/` or
        `image_encoder/` leaves that component None: `__call__` then needs `image_latents=` / `image_embeddings=`).
        `torch_dtype` is accepted and ignored: storage is bf16, accumulation fp32, on every path."""
        import json
        import os
        from .clip import CLIPVisionModelWithProjection
        from .vae import AutoencoderKLTemporalDecoder
        root = pretrained_model_name_or_path
        if not os.path.isdir(root):
            raise FileNotFoundError(f"pipeline directory not found: {root}")
        if unet is None:
            unet = UNetSpatioTemporalConditionControlNetModel.from_pretrained(root, subfolder="unet", variant=variant, device=device)
        dev = unet.device
        if controlnet is None:
            if not os.path.isdir(os.path.join(root, "controlnet")):
                raise ValueError("pass controlnet= (the SVD pipeline directory has no controlnet/ sub-folder)")
            controlnet = ControlNetSDVModel.from_pretrained(root, subfolder="controlnet", variant=variant, device=dev)
        if vae is None and os.path.isdir(os.path.join(root, "vae")):
            vae = AutoencoderKLTemporalDecoder.from_pretrained(root, subfolder="vae", variant=variant, device=dev)
        if image_encoder is None and os.path.isdir(os.path.join(root, "image_encoder")):
            image_encoder = CLIPVisionModelWithProjection.from_pretrained(root, subfolder="image_encoder", variant=variant, device=dev)
        if scheduler is None:
            kw = {}
            cfg_path = os.path.join(root, "scheduler", "scheduler_config.json")
            if os.path.exists(cfg_path):
                import inspect
                known = set(inspect.signature(EulerDiscreteScheduler.__init__).parameters) - {"self"}
                with open(cfg_path) as f:
                    kw = {k: v for k, v in json.load(f).items() if k in known}
            scheduler = EulerDiscreteScheduler(**kw)
        return cls(vae=vae, image_encoder=image_encoder, unet=unet, controlnet=controlnet, scheduler=scheduler,
                   feature_extractor=feature_extractor)

    # The memory / attention toggles the reference scripts call (run_inference_vipseg_json_repro.py:339-341).  The whole
    # model set is 4.6 GB of bf16 weights resident in 180 GB of HBM and attention is this library's own kernel, so they
    # have nothing to do; they exist so that the scripts run unchanged.
    def enable_model_cpu_offload(self, gpu_id: Optional[int] = None, device=None) -> None:
        return None

    def enable_sequential_cpu_offload(self, gpu_id: Optional[int] = None, device=None) -> None:
        return None

    def enable_xformers_memory_efficient_attention(self, attention_op=None) -> None:
        return None

    def to(self, *args, **kwargs):
        return self

    def enable_cfg_split(self, rank: int, group=None) -> None:
        """Run one branch of the CFG pair per GPU (2 ranks of `group`, rank == row: 0 uncond, 1 cond)."""
        if rank not in (0, 1):
            raise ValueError("CFG split runs on exactly 2 ranks")
        self._cfg_rows, self._cfg_group = (rank, 1), group
        self._engines.clear()

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def _execution_device(self):
        return self.unet.device

    def check_inputs(self, image, height, width):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_frames, num_channels_latents // 2, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective"
                             f" batch size of {batch_size}. Make sure the batch size matches the length of the generators.")
        if latents is None:
            gdev = generator.device if isinstance(generator, torch.Generator) else device
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def prepare_controlnet_condition(self, controlnet_condition, height: int, width: int) -> torch.Tensor:
        """pipeline...controlnet.py:500-503: the trajectory maps — a list of PIL images / numpy arrays as the reference
        scripts pass them (`control_images[:14]`), or an already pre-processed [F, 3, H, W] / [1|2, F, 3, H, W] tensor in
        [-1, 1] (e.g. from posetraj_b200.trajectory.rasterize_tracks) — become [2, F, 3, H, W] for the CFG pair."""
        cond = controlnet_condition
        if cond is None:
            raise ValueError("controlnet_condition is required")
        if not torch.is_tensor(cond):
            cond = self.image_processor.preprocess(cond, height=height, width=width)
        if cond.dim() == 4:
            cond = cond.unsqueeze(0)
        if cond.dim() != 5 or cond.shape[2] != 3:
            raise ValueError(f"controlnet_condition must be [F, 3, H, W] or [B, F, 3, H, W], got {tuple(cond.shape)}")
        if cond.shape[0] == 1:
            cond = torch.cat([cond] * 2)
        return cond

    def enable_frame_sharding(self, rank: int, world: int, group=None) -> None:
        """Shard ONE video over `world` GPUs by frames (spatial layers) / pixels (temporal layers) with an all-to-all
        around every temporal sub-block (posetraj_b200/frame_sharding.py)."""
        if not (0 <= rank < world):
            raise ValueError("rank out of range")
        self._frame_shard = (rank, world, group)
        self._engines.clear()

    def engine_for(self, frames, h, w, cond_hw):
        if self._frame_shard is not None:
            from .frame_sharding import FrameShardedEngine
            rank, world, group = self._frame_shard
            key = (frames, h, w, tuple(cond_hw), "frames", rank, world)
            if key not in self._engines:
                self._engines[key] = FrameShardedEngine(self.unet, self.controlnet, self.scheduler, frames=frames, h=h, w=w,
                                                        cond_hw=cond_hw, device=self._execution_device, rank=rank,
                                                        world=world, group=group)
            return self._engines[key]
        key = (frames, h, w, tuple(cond_hw), self._cfg_rows)
        if key not in self._engines:
            self._engines[key] = DenoiseEngine(self.unet, self.controlnet, self.scheduler, frames=frames, h=h, w=w,
                                               cond_hw=cond_hw, device=self._execution_device, rows=self._cfg_rows,
                                               group=self._cfg_group)
        return self._engines[key]

    @torch.no_grad()
    def __call__(self, image=None, controlnet_condition=None, camera_cond=None, height: int = 576, width: int = 1024,
                 num_frames: Optional[int] = None, num_inference_steps: int = 25, min_guidance_scale: float = 1.0,
                 max_guidance_scale: float = 3.0, fps: int = 7, motion_bucket_id: int = 127,
                 noise_aug_strength: float = 0.02, decode_chunk_size: Optional[int] = None,
                 num_videos_per_prompt: Optional[int] = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil", callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], return_dict: bool = True,
                 controlnet_cond_scale=1.0, batch_size=1, image_embeddings: Optional[torch.Tensor] = None,
                 image_latents: Optional[torch.Tensor] = None):
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        self.check_inputs(image, height, width)
        device = self._execution_device
        if batch_size * num_videos_per_prompt != 1:
            # the reference scripts call the pipeline once per video; batching videos into one call would change
            # results through the temporal cross-attention interleave
            raise ValueError("posetraj_b200: one video per call (batch_size * num_videos_per_prompt must be 1)")
        if max_guidance_scale <= 1.0:
            raise ValueError("posetraj_b200: the path is built for classifier-free guidance (max_guidance_scale > 1)")
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor

        # 3./4. image conditioning (outside the hot path)
        if image_embeddings is None:
            if self.image_encoder is None:
                raise ValueError("pass image_embeddings= (or construct the pipeline with an image_encoder module)")
            image_embeddings = self._encode_image(image, device, num_videos_per_prompt, True)
        if image_latents is None:
            if self.vae is None:
                raise ValueError("pass image_latents= (or construct the pipeline with a vae module)")
            # 4. encode the (noise-augmented) conditioning image (:449-461)
            img = self.image_processor.preprocess(image, height=height, width=width).to(device)
            gdev = generator.device if isinstance(generator, torch.Generator) else device
            noise = torch.randn(img.shape, generator=generator, device=gdev, dtype=img.dtype).to(device)
            img = img + noise_aug_strength * noise
            image_latents = self._encode_vae_image(img, device, num_videos_per_prompt, True)
        image_embeddings = image_embeddings.to(device=device, dtype=F32)
        image_latents = image_latents.to(device=device, dtype=F32)
        if image_latents.dim() == 4:  # [2, C, h, w] -> repeat per frame (:466)
            image_latents = image_latents.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)

        # 4. timesteps, 5. latents
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = self.scheduler.timesteps
        latents = self.prepare_latents(1, num_frames, self.unet.config.in_channels, height, width, F32, device,
                                       generator, latents)
        # controlnet condition: [F, 3, H, W] in [-1, 1] -> duplicated for the CFG pair (:500-503)
        cond = self.prepare_controlnet_condition(controlnet_condition, height, width).to(device=device, dtype=F32)
        cam = None
        if camera_cond is not None:
            cam = torch.as_tensor(camera_cond).to(device=device, dtype=F32)
            if cam.dim() == 2:
                cam = cam.unsqueeze(0)
            if cam.shape[0] == 1:
                cam = torch.cat([cam] * 2)  # ..._cam.py:506-509
        # 7. guidance scale per frame (:506-509)
        guidance = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames, device=device, dtype=F32)
        self._guidance_scale = guidance.view(1, num_frames, 1, 1, 1)
        # added_time_ids hard override [6, 128, 0.02] x2 (:513-523): the caller's fps / motion_bucket_id are ignored
        added_time_ids = _get_add_time_ids(0.02, F32, 1, 6, 128, unet=self.unet)
        added_time_ids = torch.cat([added_time_ids] * 2).to(device)

        # 8. denoising loop
        eng = self.engine_for(num_frames, h, w, tuple(cond.shape[-2:]))
        eng.load(latents=latents, image_latents=image_latents, image_embeddings=image_embeddings,
                 added_time_ids=added_time_ids, guidance=guidance, sigmas=self.scheduler.sigmas,
                 controlnet_condition=cond, camera_cond=cam, cond_scale=float(controlnet_cond_scale))
        self._num_timesteps = len(timesteps)
        use_graph = self.use_cuda_graph and callback_on_step_end is None
        for i, t in enumerate(timesteps):
            if use_graph and i == 1:
                eng.capture()
            eng.step(use_graph=use_graph and i >= 1)
            if callback_on_step_end is not None:
                if self._frame_shard is not None:
                    raise ValueError("posetraj_b200: step-end callbacks are not available with frame sharding")
                cb_latents = eng.latents.view(1, num_frames, -1, h, w)
                out = callback_on_step_end(self, i, t, {"latents": cb_latents})
                new = out.pop("latents", cb_latents) if isinstance(out, dict) else cb_latents
                if new.data_ptr() != eng.latents.data_ptr():
                    eng.latents.copy_(new.reshape(eng.latents.shape))
                # the fused update kernel has already written the NEXT step's model input from the pre-callback
                # latents; the reference re-derives latent_model_input from `latents` every step (:532-537), so rebuild
                # it (mode 1: scale by sigma[step_index], which StepAdvance has already moved on) whether the callback
                # returned a new tensor or edited the view in place
                if i + 1 < len(timesteps):
                    eng.prepare_op.launch(torch.cuda.current_stream().cuda_stream)
        if self._frame_shard is not None:
            latents = eng.gather_latents().view(1, num_frames, -1, h, w)
        else:
            latents = eng.latents.view(1, num_frames, -1, h, w).clone()

        if output_type != "latent":
            if self.vae is None:
                raise ValueError("output_type other than 'latent' needs a VAE (SURVEY.md §8f row 2)")
            frames = self.decode_latents(latents, num_frames, decode_chunk_size or num_frames)
            frames = tensor2vid(frames, self.image_processor, output_type=output_type)
        else:
            frames = latents
        if not return_dict:
            return frames
        return StableVideoDiffusionPipelineOutput(frames=frames)

    # ---- either side of the loop ------------------------------------------------------------------------
    def _encode_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        """pipeline...controlnet.py:145-172: PIL / numpy images become [0, 1] tensors (`pil_to_numpy` + `numpy_to_pt`;
        tensors are taken as they are), `_resize_with_antialiasing` to 224x224 (:602-712), CLIP `image_embeds`.  Like the
        reference, neither the [-1, 1] mapping nor CLIP's mean/std normalisation is applied here.  With
        posetraj_b200.clip.CLIPVisionModelWithProjection the resize writes the patch rows of the tower directly."""
        from .clip import resize_with_antialiasing
        if torch.is_tensor(image):
            x = image if image.dim() == 4 else image.unsqueeze(0)
        else:
            x = (self.image_processor.preprocess(image) + 1.0) / 2.0   # == numpy_to_pt(pil_to_numpy(image))
        x = x.to(device=device, dtype=F32)
        if hasattr(self.image_encoder, "encode_image"):
            emb = self.image_encoder.encode_image(x)
        else:
            emb = self.image_encoder(resize_with_antialiasing(x, (224, 224))).image_embeds.unsqueeze(1)
        emb = emb.repeat(1, num_videos_per_prompt, 1).view(emb.shape[0] * num_videos_per_prompt, 1, -1)
        if do_classifier_free_guidance:
            emb = torch.cat([torch.zeros_like(emb), emb])
        return emb

    def _encode_vae_image(self, image: torch.Tensor, device, num_videos_per_prompt, do_classifier_free_guidance):
        """pipeline...controlnet.py:174-195: mode of the VAE posterior, zeros for the unconditional branch."""
        image_latents = self.vae.encode(image.to(device=device)).latent_dist.mode()
        if do_classifier_free_guidance:
            image_latents = torch.cat([torch.zeros_like(image_latents), image_latents])
        return image_latents.repeat(num_videos_per_prompt, 1, 1, 1)

    def decode_latents(self, latents, num_frames, decode_chunk_size=14):
        """pipeline...controlnet.py:225-251: [B, F, C, h, w] latents -> [B, 3, F, H, W] fp32 frames, decoded
        `decode_chunk_size` frames at a time (the temporal layers only see the frames of one chunk, as in the
        reference)."""
        latents = latents.flatten(0, 1)
        latents = 1 / self.vae.config.scaling_factor * latents
        frames = []
        for i in range(0, latents.shape[0], decode_chunk_size):
            chunk = latents[i: i + decode_chunk_size]
            frames.append(self.vae.decode(chunk, num_frames=chunk.shape[0]).sample)
        frames = torch.cat(frames, dim=0)
        frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
        return frames.float()
