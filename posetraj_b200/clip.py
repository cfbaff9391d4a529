"""Image-conditioning branch of the pipeline on the sm_100a kernel library:

  `_encode_image`              /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:145-172
  `_resize_with_antialiasing`  /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:602-712
  image_encoder                `transformers.CLIPVisionModelWithProjection` (SVD ships ViT-H/14)

`CLIPVisionModelWithProjection` below mirrors the HF class the reference loads (same state-dict key tree, `config`
attributes, `model(pixel_values).image_embeds`), and `resize_with_antialiasing` mirrors the reference's resize helper.
Lowering: the blur is two `pt_blur_reflect` passes, the bicubic resample writes the im2col rows of the 14x14 / stride-14
patch embedding directly (`pt_bicubic_resize`), the patch embedding, every projection and MLP layer is a `pt_gemm`
(GELU in the epilogue), LayerNorm is `pt_layernorm`, attention `pt_attention_small`.  One image per call (the
reference encodes one conditioning image per video).  No CPU path.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import ops
from .engine import BF16, F32, NetPlan, Pool, WeightStore, _pad64


@dataclass
class CLIPVisionConfig:
    hidden_size: int = 1280
    intermediate_size: int = 5120
    num_hidden_layers: int = 32
    num_attention_heads: int = 16
    image_size: int = 224
    patch_size: int = 14
    projection_dim: int = 1024
    hidden_act: str = "gelu"
    layer_norm_eps: float = 1e-5
    num_channels: int = 3

    def __post_init__(self):
        if self.hidden_size % 64 or self.intermediate_size % 64:
            raise ValueError("CLIP hidden / intermediate sizes must be multiples of 64")
        if self.hidden_act not in ("gelu", "quick_gelu"):
            raise ValueError("hidden_act must be 'gelu' or 'quick_gelu'")
        if self.image_size % self.patch_size or self.hidden_size % self.num_attention_heads:
            raise ValueError("patch_size must divide image_size and heads must divide hidden_size")


def clip_param_shapes(cfg: CLIPVisionConfig) -> Dict[str, tuple]:
    d, n = cfg.hidden_size, (cfg.image_size // cfg.patch_size) ** 2
    sh: Dict[str, tuple] = {
        "vision_model.embeddings.class_embedding": (d,),
        "vision_model.embeddings.patch_embedding.weight": (d, cfg.num_channels, cfg.patch_size, cfg.patch_size),
        "vision_model.embeddings.position_embedding.weight": (n + 1, d),
        "vision_model.pre_layrnorm.weight": (d,), "vision_model.pre_layrnorm.bias": (d,),
        "vision_model.post_layernorm.weight": (d,), "vision_model.post_layernorm.bias": (d,),
        "visual_projection.weight": (cfg.projection_dim, d),
    }
    for i in range(cfg.num_hidden_layers):
        p = f"vision_model.encoder.layers.{i}."
        for nme in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sh[p + f"self_attn.{nme}.weight"] = (d, d)
            sh[p + f"self_attn.{nme}.bias"] = (d,)
        for nme in ("layer_norm1", "layer_norm2"):
            sh[p + nme + ".weight"] = (d,)
            sh[p + nme + ".bias"] = (d,)
        sh[p + "mlp.fc1.weight"], sh[p + "mlp.fc1.bias"] = (cfg.intermediate_size, d), (cfg.intermediate_size,)
        sh[p + "mlp.fc2.weight"], sh[p + "mlp.fc2.bias"] = (d, cfg.intermediate_size), (d,)
    return sh


def _gaussian_window(window_size: int, sigma: float) -> torch.Tensor:
    """`_gaussian` (:686-697), fp32 like the reference (parameter preparation on the host: <= a few dozen taps)."""
    s = torch.tensor([[sigma]], dtype=F32)
    x = (torch.arange(window_size, dtype=F32) - window_size // 2).expand(1, -1)
    if window_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * s.pow(2.0)))
    return (g / g.sum(-1, keepdim=True))[0].contiguous()


class ResizePlan:
    """`_resize_with_antialiasing(image, (S, S))` for one [C, H, W] fp32 image on the device."""

    def __init__(self, C: int, H: int, W: int, size: int, device, *, patches: Optional[torch.Tensor] = None, patch: int = 1,
                 want_f32: bool = True):
        factors = (H / size, W / size)
        sig = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
        ks = [int(max(2.0 * 2 * sig[0], 3)), int(max(2.0 * 2 * sig[1], 3))]
        ks = [k + 1 if k % 2 == 0 else k for k in ks]
        self.wy = _gaussian_window(ks[0], sig[0]).to(device)
        self.wx = _gaussian_window(ks[1], sig[1]).to(device)
        self.x_in = torch.zeros(C, H, W, device=device, dtype=F32)
        self.tmp = torch.zeros_like(self.x_in)
        self.blur = torch.zeros_like(self.x_in)
        self.out = torch.zeros(C, size, size, device=device, dtype=F32) if want_f32 else None
        self.ops: List = [ops.BlurReflect(self.x_in, self.tmp, self.wx, axis=0, name="resize.blur_x"),
                          ops.BlurReflect(self.tmp, self.blur, self.wy, axis=1, name="resize.blur_y"),
                          ops.BicubicResize(self.blur, size=size, out_f32=self.out, out_patches=patches, patch=patch,
                                            name="resize.bicubic")]


def resize_with_antialiasing(image: torch.Tensor, size=(224, 224)) -> torch.Tensor:
    """Mirror of the reference helper (:602-632) for CUDA tensors [N, C, H, W] or [C, H, W]; square targets."""
    if image.device.type != "cuda":
        raise RuntimeError("posetraj_b200: inputs must be CUDA tensors (no CPU fallback)")
    if size[0] != size[1]:
        raise ValueError("posetraj_b200.resize_with_antialiasing: square target sizes only")
    if image.dim() == 3:
        image = image.unsqueeze(0)
    sp = torch.cuda.current_stream().cuda_stream
    outs = []
    for img in image:
        plan = ResizePlan(img.shape[0], img.shape[1], img.shape[2], size[0], image.device)
        plan.x_in.copy_(img)
        NetPlan.run(plan.ops, sp)
        outs.append(plan.out)
    return torch.stack(outs, 0)


class ClipPlan:
    """One image through the vision tower: [patch rows] -> image_embeds [1, projection_dim]."""

    def __init__(self, cfg: CLIPVisionConfig, w: WeightStore, device):
        d, heads = cfg.hidden_size, cfg.num_attention_heads
        pp = cfg.image_size // cfg.patch_size
        n_patch, S = pp * pp, pp * pp + 1
        kp = cfg.num_channels * cfg.patch_size ** 2
        self.kpad = _pad64(kp)
        act = 2 if cfg.hidden_act == "gelu" else 3
        eps = cfg.layer_norm_eps
        self.pool = Pool(device)
        self.ops: List = []
        self.patches = torch.zeros(n_patch, self.kpad, device=device, dtype=BF16)   # pad columns stay zero
        # patch embedding as a GEMM over im2col rows; position embeddings of the patch tokens as the residual operand
        w_patch = torch.zeros(d, self.kpad, device=device, dtype=BF16)
        w_patch[:, :kp] = w._get("vision_model.embeddings.patch_embedding.weight").to(F32).reshape(d, kp).to(BF16)
        pos = w._get("vision_model.embeddings.position_embedding.weight").to(F32)
        self.pos_patch = pos[1:].to(BF16).contiguous()
        self.h0 = torch.zeros(S, d, device=device, dtype=BF16)
        self.h0[0] = (w._get("vision_model.embeddings.class_embedding").to(F32) + pos[0]).to(BF16)   # constant row
        self._keep = [w_patch]
        self.ops.append(ops.Gemm(self.patches, w_patch, self.h0[1:], res1=self.pos_patch, alg_k=kp,
                                 name="vision_model.embeddings.patch_embedding"))

        def ln(x, key):
            out = self.pool.get(x.shape[0], x.shape[1])
            self.ops.append(ops.LayerNorm(x, out, w.f32(key + ".weight"), w.f32(key + ".bias"), eps=eps, name=key))
            return out

        def gemm(a0, wt, n, **kw):
            out = self.pool.get(a0.shape[0], n)
            self.ops.append(ops.Gemm(a0, wt, out, **kw))
            return out

        x = ln(self.h0, "vision_model.pre_layrnorm")
        for i in range(cfg.num_hidden_layers):
            p = f"vision_model.encoder.layers.{i}."
            l1 = ln(x, p + "layer_norm1")
            wqkv = w.cat_rows([p + f"self_attn.{n}_proj.weight" for n in "qkv"], "bf16")
            bqkv = w.cat_rows([p + f"self_attn.{n}_proj.bias" for n in "qkv"], "f32")
            qkv = gemm(l1, wqkv, 3 * d, bias=bqkv, name=p + "self_attn.qkv")
            self.pool.put(l1)
            att = self.pool.get(S, d)
            self.ops.append(ops.AttnSmall(qkv, att, heads=heads, name=p + "self_attn"))
            self.pool.put(qkv)
            x2 = gemm(att, w.linear(p + "self_attn.out_proj.weight"), d, bias=w.f32(p + "self_attn.out_proj.bias"), res1=x,
                      name=p + "self_attn.out_proj")
            self.pool.put(att, x)
            l2 = ln(x2, p + "layer_norm2")
            f1 = gemm(l2, w.linear(p + "mlp.fc1.weight"), cfg.intermediate_size, bias=w.f32(p + "mlp.fc1.bias"),
                      act_silu=act, name=p + "mlp.fc1")
            self.pool.put(l2)
            x = gemm(f1, w.linear(p + "mlp.fc2.weight"), d, bias=w.f32(p + "mlp.fc2.bias"), res1=x2, name=p + "mlp.fc2")
            self.pool.put(f1, x2)
        pooled = torch.zeros(1, d, device=device, dtype=BF16)
        self.ops.append(ops.LayerNorm(x[0:1], pooled, w.f32("vision_model.post_layernorm.weight"),
                                      w.f32("vision_model.post_layernorm.bias"), eps=eps, name="vision_model.post_layernorm"))
        self.embeds = torch.zeros(1, cfg.projection_dim, device=device, dtype=F32)
        self.ops.append(ops.Gemm(pooled, w.linear("visual_projection.weight"), self.embeds, name="visual_projection"))


def load_clip_checkpoint(path: str, subfolder: Optional[str] = None, variant: Optional[str] = None, **overrides):
    """(CLIPVisionConfig, state dict) from a Hugging Face `CLIPVisionModelWithProjection` directory (`config.json` +
    `model[.variant].safetensors` or `pytorch_model[.variant].bin`), the layout of SVD's `image_encoder/` subfolder the
    reference loads (scripts/run_inference_vipseg_json_repro.py:335-339).  Host only; no transformers import."""
    import json
    import os
    from dataclasses import fields
    from .checkpoint import resolve_dir
    d = resolve_dir(path, subfolder)
    known = {f.name for f in fields(CLIPVisionConfig)}
    ckw: Dict = {}
    cp = os.path.join(d, "config.json")
    if os.path.exists(cp):
        with open(cp) as f:
            for k, v in json.load(f).items():
                if k in known and v is not None:
                    ckw[k] = v
    ckw.update({k: v for k, v in overrides.items() if k in known})
    v = f".{variant}" if variant else ""
    for name in (f"model{v}.safetensors", f"pytorch_model{v}.bin"):
        fp = os.path.join(d, name)
        if os.path.exists(fp):
            if fp.endswith(".safetensors"):
                from safetensors.torch import load_file
                sd = load_file(fp, device="cpu")
            else:
                sd = torch.load(fp, map_location="cpu", weights_only=True)
            return CLIPVisionConfig(**ckw), sd
    raise FileNotFoundError(f"no model{v}.safetensors / pytorch_model{v}.bin in {d}")


@dataclass
class CLIPVisionModelOutput:
    image_embeds: torch.Tensor


class CLIPVisionModelWithProjection(torch.nn.Module):
    def __init__(self, cfg: Optional[CLIPVisionConfig] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 device=None):
        super().__init__()
        cfg = cfg or CLIPVisionConfig()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("posetraj_b200 runs on CUDA sm_100a only; there is no CPU path")
        if state_dict is None:
            raise ValueError("state_dict is required (use from_random / from_pretrained)")
        exp = clip_param_shapes(cfg)
        missing = [k for k in exp if k not in state_dict]
        if missing:
            raise KeyError(f"CLIPVisionModelWithProjection: state dict misses {len(missing)} keys, e.g. {missing[:3]}")
        for k, shp in exp.items():
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {shp}, got {tuple(state_dict[k].shape)}")
        self.cfg = cfg
        self._device = torch.device(device)
        self._sd = {k: state_dict[k] for k in exp}   # HF checkpoints also carry position_ids buffers: ignored
        self.weights = WeightStore(self._sd, self._device)
        self.config = SimpleNamespace(**asdict(cfg))
        self.dtype = BF16
        self._plan: Optional[ClipPlan] = None
        self._resize: Dict[tuple, ResizePlan] = {}

    @property
    def device(self):
        return self._device

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def parameters(self, recurse: bool = True):   # the reference reads `next(image_encoder.parameters()).dtype`
        return iter([torch.nn.Parameter(torch.zeros(1, device=self._device, dtype=BF16), requires_grad=False)])

    def num_parameters(self) -> int:
        return sum(int(math.prod(s)) for s in clip_param_shapes(self.cfg).values())

    @classmethod
    def from_random(cls, cfg: Optional[CLIPVisionConfig] = None, device=None, seed: int = 0):
        cfg = cfg or CLIPVisionConfig()
        device = device or torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=device).manual_seed(seed)
        sd = {}
        for k, shp in clip_param_shapes(cfg).items():
            if "norm" in k and k.endswith("weight"):
                sd[k] = torch.ones(shp, device=device, dtype=F32)
            elif "norm" in k or k.endswith("bias"):
                sd[k] = torch.zeros(shp, device=device, dtype=F32)
            else:
                std = 0.02 if len(shp) > 1 else cfg.hidden_size ** -0.5
                sd[k] = (torch.randn(shp, device=device, generator=g, dtype=F32) * std).to(BF16 if len(shp) > 1 else F32)
        return cls(cfg, sd, device)

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, variant: Optional[str] = None, device=None, **kw):
        cfg, sd = load_clip_checkpoint(path, subfolder, variant, **kw)
        return cls(cfg, sd, device)

    def _plan_for(self) -> ClipPlan:
        if self._plan is None:
            self._plan = ClipPlan(self.cfg, self.weights, self._device)
        return self._plan

    def forward(self, pixel_values: torch.Tensor, return_dict: bool = True):
        """`image_encoder(image).image_embeds` for [N, 3, image_size, image_size] pixel values."""
        if pixel_values.device.type != "cuda":
            raise RuntimeError("posetraj_b200: inputs must be CUDA tensors (no CPU fallback)")
        S = self.cfg.image_size
        if pixel_values.dim() != 4 or tuple(pixel_values.shape[1:]) != (self.cfg.num_channels, S, S):
            raise ValueError(f"expected [N, {self.cfg.num_channels}, {S}, {S}], got {tuple(pixel_values.shape)}")
        plan = self._plan_for()
        sp = torch.cuda.current_stream().cuda_stream
        P, pp = self.cfg.patch_size, S // self.cfg.patch_size
        kp = self.cfg.num_channels * P * P
        outs = []
        for img in pixel_values.to(F32):
            # im2col of the already-resized image: a layout change of the input, [C, pp, P, pp, P] -> [pp*pp, C*P*P]
            rows = img.view(-1, pp, P, pp, P).permute(1, 3, 0, 2, 4).reshape(pp * pp, kp)
            plan.patches[:, :kp].copy_(rows)
            NetPlan.run(plan.ops, sp)
            outs.append(plan.embeds.clone())
        out = torch.cat(outs, 0)
        return CLIPVisionModelOutput(image_embeds=out) if return_dict else (out,)

    def encode_image(self, image01: torch.Tensor) -> torch.Tensor:
        """`_encode_image` (:145-172) up to the CFG duplication, fused: image in [0, 1] ([N, 3, H, W], any size) ->
        anti-aliased resize written straight into the patch rows -> vision tower -> [N, 1, projection_dim]."""
        if image01.device.type != "cuda":
            raise RuntimeError("posetraj_b200: inputs must be CUDA tensors (no CPU fallback)")
        if image01.dim() == 3:
            image01 = image01.unsqueeze(0)
        plan = self._plan_for()
        sp = torch.cuda.current_stream().cuda_stream
        outs = []
        for img in image01.to(F32):
            key = tuple(img.shape)
            if key not in self._resize:
                self._resize[key] = ResizePlan(img.shape[0], img.shape[1], img.shape[2], self.cfg.image_size, self._device,
                                               patches=plan.patches, patch=self.cfg.patch_size, want_f32=False)
            rp = self._resize[key]
            rp.x_in.copy_(img)
            NetPlan.run(rp.ops, sp)
            NetPlan.run(plan.ops, sp)
            outs.append(plan.embeds.clone())
        return torch.cat(outs, 0).unsqueeze(1)
