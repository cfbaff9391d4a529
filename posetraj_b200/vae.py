"""Mirror of diffusers' `AutoencoderKLTemporalDecoder` (the SVD VAE) on the sm_100a kernel library — SURVEY.md §8f
row 2, the steps on either side of the denoise loop:

  `_encode_vae_image`  /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:174-195
                       (`vae.encode(image).latent_dist.mode()`)
  `decode_latents`     /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:225-251
                       (`vae.decode(latents, num_frames=chunk).sample`)

Same call surface as the diffusers class the reference uses (`encode(x).latent_dist.mode()`, `decode(z,
num_frames=...).sample`, `config.scaling_factor`, `config.block_out_channels`, `config.force_upcast`, `dtype`) and the
diffusers state-dict key tree (`encoder.*`, `decoder.*`, `quant_conv.*`), so a real SVD VAE checkpoint loads as is.

Lowering (same kernels as the denoise step; nothing here is a new hot loop):
  * every 3x3 conv, 1x1 shortcut, temporal (3,1,1) conv and attention projection is a `pt_gemm` launch on token-major
    bf16 rows (3x3: zero-haloed rows, a tap is a row shift); the encoder's `Downsample2D(padding=0)` — pad right/bottom
    by one, stride 2 — is the same implicit GEMM with taps (0..2, 0..2) instead of (-1..1, -1..1): the halo IS the pad;
  * GroupNorm(+SiLU) is `pt_groupnorm` (4-D statistics per image, 5-D per video for TemporalResnetBlock);
  * the mid-block attention has ONE head of dimension C = 512, which is a plain GEMM, not a flash tile: per image
    S = q k^T / sqrt(C) (fp32, `pt_gemm`), `pt_softmax_rows`, O = P v (`pt_gemm` against v^T, which is produced
    directly by swapping the operands of the v projection; v's bias is added after P v since the rows of P sum to 1);
  * the AlphaBlender (merge_strategy "learned", switch_spatial_to_temporal_mix) of x and x + f(x) is
    x + sigmoid(mix_factor) f(x): an epilogue scale of the last temporal conv;
  * `time_conv_out` + the NCHW fp32 hand-off is `pt_time_conv3`.
No CPU path: construction on a non-CUDA device raises.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .engine import BF16, F32, NetPlan, Pool, WeightStore, _pad64


@dataclass
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    scaling_factor: float = 0.18215
    force_upcast: bool = True

    def __post_init__(self):
        self.block_out_channels = tuple(self.block_out_channels)
        for c in self.block_out_channels:
            if c % 64:
                raise ValueError("VAE block_out_channels must be multiples of 64")


# ---------------------------------------------------------------------------------------------------------------
# state-dict key tree (diffusers naming)
# ---------------------------------------------------------------------------------------------------------------
def _conv(sh: Dict, p: str, cin: int, cout: int, k: int = 3) -> None:
    sh[p + ".weight"] = (cout, cin, k, k)
    sh[p + ".bias"] = (cout,)


def _norm(sh: Dict, p: str, c: int) -> None:
    sh[p + ".weight"] = (c,)
    sh[p + ".bias"] = (c,)


def _resnet2d(sh: Dict, p: str, cin: int, cout: int) -> None:
    _norm(sh, p + "norm1", cin)
    _conv(sh, p + "conv1", cin, cout)
    _norm(sh, p + "norm2", cout)
    _conv(sh, p + "conv2", cout, cout)
    if cin != cout:
        _conv(sh, p + "conv_shortcut", cin, cout, 1)


def _st_resblock(sh: Dict, p: str, cin: int, cout: int) -> None:
    _resnet2d(sh, p + "spatial_res_block.", cin, cout)
    t = p + "temporal_res_block."
    for n in ("1", "2"):
        _norm(sh, t + "norm" + n, cout)
        sh[t + f"conv{n}.weight"] = (cout, cout, 3, 1, 1)
        sh[t + f"conv{n}.bias"] = (cout,)
    sh[p + "time_mixer.mix_factor"] = (1,)


def _attention(sh: Dict, p: str, c: int) -> None:
    _norm(sh, p + "group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        sh[p + n + ".weight"] = (c, c)
        sh[p + n + ".bias"] = (c,)


def vae_param_shapes(cfg: VaeConfig) -> Dict[str, tuple]:
    sh: Dict[str, tuple] = {}
    ch = list(cfg.block_out_channels)
    # encoder
    _conv(sh, "encoder.conv_in", cfg.in_channels, ch[0])
    prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet2d(sh, f"encoder.down_blocks.{i}.resnets.{j}.", prev if j == 0 else c, c)
        if i < len(ch) - 1:
            _conv(sh, f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c)
        prev = c
    _resnet2d(sh, "encoder.mid_block.resnets.0.", ch[-1], ch[-1])
    _attention(sh, "encoder.mid_block.attentions.0.", ch[-1])
    _resnet2d(sh, "encoder.mid_block.resnets.1.", ch[-1], ch[-1])
    _norm(sh, "encoder.conv_norm_out", ch[-1])
    _conv(sh, "encoder.conv_out", ch[-1], 2 * cfg.latent_channels)
    # temporal decoder
    _conv(sh, "decoder.conv_in", cfg.latent_channels, ch[-1])
    for j in range(cfg.layers_per_block):
        _st_resblock(sh, f"decoder.mid_block.resnets.{j}.", ch[-1], ch[-1])
    _attention(sh, "decoder.mid_block.attentions.0.", ch[-1])
    rev = list(reversed(ch))
    prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _st_resblock(sh, f"decoder.up_blocks.{i}.resnets.{j}.", prev if j == 0 else c, c)
        if i < len(rev) - 1:
            _conv(sh, f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c)
        prev = c
    _norm(sh, "decoder.conv_norm_out", ch[0])
    _conv(sh, "decoder.conv_out", ch[0], cfg.out_channels)
    sh["decoder.time_conv_out.weight"] = (cfg.out_channels, cfg.out_channels, 3, 1, 1)
    sh["decoder.time_conv_out.bias"] = (cfg.out_channels,)
    _conv(sh, "quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    return sh


# ---------------------------------------------------------------------------------------------------------------
# lowering
# ---------------------------------------------------------------------------------------------------------------
class _VaePlan:
    """Shared block builders: an in-order op list over pooled token-major bf16 buffers."""

    def __init__(self, weights: WeightStore, device, n_img: int):
        self.w, self.device, self.n = weights, device, n_img
        self.pool = Pool(device)
        self.ops: List = []
        self.stats = torch.zeros((2 * max(n_img, 1) + 4 * ops.NUM_SMS + 64) * 64 + 1024, device=device, dtype=torch.float64)

    # -- primitives ---------------------------------------------------------------------------------------
    def gn(self, x, key, *, rows_per_stat, eps, silu, halo=None):
        Cc = x.shape[1]
        if halo is not None:
            n_img = x.shape[0] // (halo[0] * halo[1])
            out = self.pool.get(n_img * (halo[0] + 1) * (halo[1] + 1), Cc)
        else:
            out = self.pool.get(x.shape[0], Cc)
        self.ops.append(ops.GroupNorm(x, out, self.w.f32(key + ".weight"), self.w.f32(key + ".bias"), self.stats,
                                      rows_per_stat=rows_per_stat, eps=eps, silu=silu, halo=halo, name=key))
        return out

    def gemm(self, a0, wt, n_cols, *, out=None, name="gemm", **kw):
        if out is None:
            out = self.pool.get(kw.pop("out_rows", a0.shape[0]), n_cols)
        else:
            kw.pop("out_rows", None)
        self.ops.append(ops.Gemm(a0, wt, out, name=name, **kw))
        return out

    # -- blocks -------------------------------------------------------------------------------------------
    def resnet2d(self, p: str, x, cout: int, hw: tuple, eps: float = 1e-6):
        """ResnetBlock2D(temb_channels=None): x + conv2(silu(norm2(conv1(silu(norm1(x)))))), 1x1 shortcut if needed."""
        w = self.w
        H, W = hw
        rows = self.n * H * W
        cin = x.shape[1]
        taps = ops.conv3x3_taps(W)
        g1 = self.gn(x, p + "norm1", rows_per_stat=H * W, eps=eps, silu=True, halo=hw)
        h1 = self.gemm(g1, w.conv3(p + "conv1.weight"), cout, taps=taps, bias=w.f32(p + "conv1.bias"), halo=hw,
                       out_rows=rows, name=p + "conv1")
        self.pool.put(g1)
        g2 = self.gn(h1, p + "norm2", rows_per_stat=H * W, eps=eps, silu=True, halo=hw)
        self.pool.put(h1)
        if cin != cout:
            sc = self.gemm(x, w.linear(p + "conv_shortcut.weight"), cout, bias=w.f32(p + "conv_shortcut.bias"),
                           name=p + "conv_shortcut")
        else:
            sc = x
        out = self.gemm(g2, w.conv3(p + "conv2.weight"), cout, taps=taps, bias=w.f32(p + "conv2.bias"), res1=sc, halo=hw,
                        out_rows=rows, name=p + "conv2")
        self.pool.put(g2)
        if sc is not x:
            self.pool.put(sc)
        return out

    def st_resblock(self, p: str, x, cout: int, hw: tuple, frames: int):
        """SpatioTemporalResBlock(temb_channels=None, merge_strategy="learned", switch_spatial_to_temporal_mix=True)."""
        w = self.w
        HW = hw[0] * hw[1]
        B = self.n // frames
        xs = self.resnet2d(p + "spatial_res_block.", x, cout, hw, eps=1e-6)
        t = p + "temporal_res_block."
        taps = (-HW, 0, HW)
        t1 = self.gn(xs, t + "norm1", rows_per_stat=frames * HW, eps=1e-5, silu=True)
        t2 = self.gemm(t1, w.tconv(t + "conv1.weight"), cout, batches=B, taps=taps, bias=w.f32(t + "conv1.bias"),
                       name=t + "conv1")
        self.pool.put(t1)
        t3 = self.gn(t2, t + "norm2", rows_per_stat=frames * HW, eps=1e-5, silu=True)
        self.pool.put(t2)
        # blend: (1 - s) xs + s (xs + f) with s = sigmoid(mix_factor)  ==  xs + s f
        s = w.alpha(p + "time_mixer.mix_factor")
        out = self.gemm(t3, w.tconv(t + "conv2.weight"), cout, batches=B, taps=taps, bias=w.f32(t + "conv2.bias"),
                        acc_scale=s, res1=xs, name=t + "conv2+mix")
        self.pool.put(t3, xs)
        return out

    def attention(self, p: str, x, hw: tuple):
        """diffusers Attention(heads=1, dim_head=C, group_norm, bias, residual_connection) on [n, C, H, W]."""
        w, dev = self.w, self.device
        HW = hw[0] * hw[1]
        Cc = x.shape[1]
        Sp = (HW + 63) // 64 * 64
        g = self.gn(x, p + "group_norm", rows_per_stat=HW, eps=1e-6, silu=False)
        wqk = w.cat_rows([p + "to_q.weight", p + "to_k.weight"], "bf16")
        bqk = w.cat_rows([p + "to_q.bias", p + "to_k.bias"], "f32")
        qk = self.gemm(g, wqk, 2 * Cc, bias=bqk, name=p + "to_qk")
        att = self.pool.get(x.shape[0], Cc)
        # per-image scratch, reused in stream order; pad columns stay zero (they are never written)
        vt = torch.zeros(Cc, Sp, device=dev, dtype=BF16)
        logits = torch.zeros(HW, Sp, device=dev, dtype=F32)
        prob = torch.zeros(HW, Sp, device=dev, dtype=BF16)
        wv, bv = w.linear(p + "to_v.weight"), w.f32(p + "to_v.bias")
        for i in range(self.n):
            rows = slice(i * HW, (i + 1) * HW)
            self.ops.append(ops.Gemm(wv, g[rows], vt, n_out=HW, name=p + f"to_v^T[{i}]"))
            self.ops.append(ops.Gemm(qk[rows, :Cc], qk[rows, Cc:], logits, n_out=HW, acc_scale=1.0 / math.sqrt(Cc),
                                     name=p + f"qk^T[{i}]"))
            self.ops.append(ops.SoftmaxRows(logits, prob, HW, name=p + f"softmax[{i}]"))
            self.ops.append(ops.Gemm(prob, vt, att[rows], bias=bv, name=p + f"pv[{i}]"))
        self._keep = getattr(self, "_keep", []) + [vt, logits, prob]
        self.pool.put(g, qk)
        out = self.gemm(att, w.linear(p + "to_out.0.weight"), Cc, bias=w.f32(p + "to_out.0.bias"), res1=x,
                        name=p + "to_out")
        self.pool.put(att)
        return out


class VaeDecodePlan(_VaePlan):
    """TemporalDecoder.forward for a fixed (videos, frames per video, latent h, w)."""

    def __init__(self, cfg: VaeConfig, weights: WeightStore, *, batch: int, frames: int, h: int, w: int, device):
        super().__init__(weights, device, batch * frames)
        ch = list(cfg.block_out_channels)
        wt = weights
        self.cfg, self.B, self.F, self.h, self.w_lat = cfg, batch, frames, h, w
        n = self.n
        self.cin_pad = _pad64(cfg.latent_channels)
        self.z_in = torch.zeros(n * (h + 1) * (w + 1), self.cin_pad, device=device, dtype=BF16)
        hw = (h, w)
        x = self.gemm(self.z_in, wt.conv3("decoder.conv_in.weight", self.cin_pad), ch[-1], taps=ops.conv3x3_taps(w),
                      bias=wt.f32("decoder.conv_in.bias"), halo=hw, out_rows=n * h * w, name="decoder.conv_in",
                      alg_k=9 * cfg.latent_channels)
        # mid block: resnet, (attention, resnet) ...
        y = self.st_resblock("decoder.mid_block.resnets.0.", x, ch[-1], hw, frames)
        self.pool.put(x)
        x = y
        for j in range(1, cfg.layers_per_block):
            if j == 1:
                y = self.attention("decoder.mid_block.attentions.0.", x, hw)
                self.pool.put(x)
                x = y
            y = self.st_resblock(f"decoder.mid_block.resnets.{j}.", x, ch[-1], hw, frames)
            self.pool.put(x)
            x = y
        rev = list(reversed(ch))
        for i, c in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                y = self.st_resblock(f"decoder.up_blocks.{i}.resnets.{j}.", x, c, hw, frames)
                self.pool.put(x)
                x = y
            if i < len(rev) - 1:
                key = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                H, W = hw
                xh = self.pool.get(n * (2 * H + 1) * (2 * W + 1), c)
                self.ops.append(ops.Upsample2x(x, xh, n=n, H=H, W=W, halo=True, scale=2, name=key + ".nearest2x"))
                self.pool.put(x)
                hw = (2 * H, 2 * W)
                x = self.gemm(xh, wt.conv3(key + ".weight"), c, taps=ops.conv3x3_taps(hw[1]), bias=wt.f32(key + ".bias"),
                              halo=hw, out_rows=n * hw[0] * hw[1], name=key)
                self.pool.put(xh)
        g = self.gn(x, "decoder.conv_norm_out", rows_per_stat=hw[0] * hw[1], eps=1e-6, silu=True, halo=hw)
        self.pool.put(x)
        # conv_out in fp32 (the frames leave the library as fp32), then the 3-tap conv over frames + NCHW hand-off
        self.rgb = torch.zeros(n * hw[0] * hw[1], 4, device=device, dtype=F32)
        self.ops.append(ops.Gemm(g, wt.conv3("decoder.conv_out.weight"), self.rgb, n_out=cfg.out_channels,
                                 taps=ops.conv3x3_taps(hw[1]), bias=wt.f32("decoder.conv_out.bias"), halo=hw, block_n=32,
                                 name="decoder.conv_out"))
        self.pool.put(g)
        self.out_hw = hw
        self.sample = torch.zeros(n, cfg.out_channels, hw[0], hw[1], device=device, dtype=F32)
        self.ops.append(ops.TimeConv3(self.rgb, wt.f32("decoder.time_conv_out.weight"), wt.f32("decoder.time_conv_out.bias"),
                                      self.sample, batch=batch, frames=frames, hw=hw[0] * hw[1],
                                      name="decoder.time_conv_out"))


class VaeEncodePlan(_VaePlan):
    """Encoder.forward + quant_conv for a fixed (images, H, W)."""

    def __init__(self, cfg: VaeConfig, weights: WeightStore, *, n_img: int, H: int, W: int, device):
        super().__init__(weights, device, n_img)
        ch = list(cfg.block_out_channels)
        wt = weights
        down = 2 ** (len(ch) - 1)
        if H % down or W % down:
            raise ValueError(f"image height/width must be divisible by {down}")
        n = n_img
        self.cin_pad = _pad64(cfg.in_channels)
        self.x_in = torch.zeros(n * (H + 1) * (W + 1), self.cin_pad, device=device, dtype=BF16)
        hw = (H, W)
        x = self.gemm(self.x_in, wt.conv3("encoder.conv_in.weight", self.cin_pad), ch[0], taps=ops.conv3x3_taps(W),
                      bias=wt.f32("encoder.conv_in.bias"), halo=hw, out_rows=n * H * W, name="encoder.conv_in",
                      alg_k=9 * cfg.in_channels)
        for i, c in enumerate(ch):
            for j in range(cfg.layers_per_block):
                y = self.resnet2d(f"encoder.down_blocks.{i}.resnets.{j}.", x, c, hw)
                self.pool.put(x)
                x = y
            if i < len(ch) - 1:
                # Downsample2D(padding=0): pad right/bottom by one (= the halo), 3x3 stride 2, taps (0..2, 0..2)
                key = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                Hc, Wc = hw
                xh = self.pool.get(n * (Hc + 1) * (Wc + 1), c)
                self.ops.append(ops.Upsample2x(x, xh, n=n, H=Hc, W=Wc, halo=True, scale=1, name=key + ".halo"))
                self.pool.put(x)
                taps = [ky * (Wc + 1) + kx for ky in range(3) for kx in range(3)]
                x = self.gemm(xh, wt.conv3(key + ".weight"), c, taps=taps, bias=wt.f32(key + ".bias"), halo=hw, ostride=2,
                              out_rows=n * (Hc // 2) * (Wc // 2), name=key)
                self.pool.put(xh)
                hw = (Hc // 2, Wc // 2)
        y = self.resnet2d("encoder.mid_block.resnets.0.", x, ch[-1], hw)
        self.pool.put(x)
        x = self.attention("encoder.mid_block.attentions.0.", y, hw)
        self.pool.put(y)
        y = self.resnet2d("encoder.mid_block.resnets.1.", x, ch[-1], hw)
        self.pool.put(x)
        g = self.gn(y, "encoder.conv_norm_out", rows_per_stat=hw[0] * hw[1], eps=1e-6, silu=True, halo=hw)
        self.pool.put(y)
        m2 = 2 * cfg.latent_channels
        self.pre = torch.zeros(n * hw[0] * hw[1], m2, device=device, dtype=F32)
        self.ops.append(ops.Gemm(g, wt.conv3("encoder.conv_out.weight"), self.pre, n_out=m2, taps=ops.conv3x3_taps(hw[1]),
                                 bias=wt.f32("encoder.conv_out.bias"), halo=hw, block_n=32, name="encoder.conv_out"))
        self.pool.put(g)
        # quant_conv (1x1, 8 -> 8) on fp32 rows
        self.moments = torch.zeros(n * hw[0] * hw[1], m2, device=device, dtype=F32)
        self.ops.append(ops.SmallLinear(self.pre, wt.linear("quant_conv.weight"), self.moments, wt.f32("quant_conv.bias"),
                                        name="quant_conv"))
        self.out_hw = hw


# ---------------------------------------------------------------------------------------------------------------
# reference-shaped module
# ---------------------------------------------------------------------------------------------------------------
class DiagonalGaussianDistribution:
    """`latent_dist` of `vae.encode(...)`: the path uses `.mode()`; `.sample()` is provided for completeness."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator=None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKLTemporalDecoder(torch.nn.Module):
    def __init__(self, cfg: Optional[VaeConfig] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None, device=None):
        super().__init__()
        cfg = cfg or VaeConfig()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("posetraj_b200 runs on CUDA sm_100a only; there is no CPU path")
        self.cfg = cfg
        self._device = torch.device(device)
        exp = vae_param_shapes(cfg)
        if state_dict is None:
            raise ValueError("state_dict is required (use from_random / from_pretrained)")
        missing = [k for k in exp if k not in state_dict]
        if missing:
            raise KeyError(f"AutoencoderKLTemporalDecoder: state dict misses {len(missing)} keys, e.g. {missing[:3]}")
        for k, shp in exp.items():
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {shp}, got {tuple(state_dict[k].shape)}")
        self._sd = dict(state_dict)
        self.weights = WeightStore(self._sd, self._device)
        self.config = SimpleNamespace(**asdict(cfg))
        self.dtype = BF16
        self._dec: Dict[tuple, VaeDecodePlan] = {}
        self._enc: Dict[tuple, VaeEncodePlan] = {}

    @property
    def device(self):
        return self._device

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def num_parameters(self) -> int:
        return sum(int(math.prod(s)) for s in vae_param_shapes(self.cfg).values())

    @classmethod
    def from_random(cls, cfg: Optional[VaeConfig] = None, device=None, seed: int = 0):
        from .models import random_state_dict
        cfg = cfg or VaeConfig()
        device = device or torch.device("cuda", torch.cuda.current_device())
        shapes = vae_param_shapes(cfg)
        sd = random_state_dict(shapes, device, seed)
        for k, shp in shapes.items():
            if k.endswith("mix_factor"):
                sd[k] = torch.zeros(1, device=device, dtype=F32)   # AlphaBlender(alpha=merge_factor=0.0)
            elif "norm" in k.split(".")[-2]:                        # GroupNorm affine: ones / zeros
                sd[k] = (torch.ones if k.endswith("weight") else torch.zeros)(shp, device=device, dtype=F32)
        return cls(cfg, sd, device)

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, variant: Optional[str] = None, device=None, **kw):
        import json
        import os
        from dataclasses import fields
        from .checkpoint import load_state_dict, resolve_dir
        d = resolve_dir(path, subfolder)
        known = {f.name for f in fields(VaeConfig)}
        ckw: Dict = {}
        cp = os.path.join(d, "config.json")
        if os.path.exists(cp):
            with open(cp) as f:
                for k, v in json.load(f).items():
                    if k in known and v is not None:
                        ckw[k] = tuple(v) if isinstance(v, list) else v
        ckw.update({k: v for k, v in kw.items() if k in known})
        return cls(VaeConfig(**ckw), load_state_dict(d, variant), device)

    # -- encode / decode --------------------------------------------------------------------------------
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        if x.device.type != "cuda":
            raise RuntimeError("posetraj_b200: inputs must be CUDA tensors (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != self.cfg.in_channels:
            raise ValueError(f"expected [N, {self.cfg.in_channels}, H, W], got {tuple(x.shape)}")
        n, _, H, W = x.shape
        key = (n, H, W)
        if key not in self._enc:
            self._enc[key] = VaeEncodePlan(self.cfg, self.weights, n_img=n, H=H, W=W, device=self._device)
        plan = self._enc[key]
        sp = torch.cuda.current_stream().cuda_stream
        xin = x if x.dtype in (F32, BF16) else x.to(F32)
        ops.Layout(xin.contiguous(), plan.x_in, to_tokens=True, halo=True).launch(sp)
        NetPlan.run(plan.ops, sp)
        h, w = plan.out_hw
        moments = plan.moments.view(n, h * w, -1).transpose(1, 2).reshape(n, -1, h, w).clone()
        dist = DiagonalGaussianDistribution(moments)
        return AutoencoderKLOutput(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z: torch.Tensor, num_frames: int = 1, return_dict: bool = True, image_only_indicator=None):
        if z.device.type != "cuda":
            raise RuntimeError("posetraj_b200: inputs must be CUDA tensors (no CPU fallback)")
        if z.dim() != 4 or z.shape[1] != self.cfg.latent_channels:
            raise ValueError(f"expected [N, {self.cfg.latent_channels}, h, w], got {tuple(z.shape)}")
        n, _, h, w = z.shape
        if num_frames < 1 or n % num_frames:
            raise ValueError("the number of latents must be a multiple of num_frames")
        key = (n // num_frames, num_frames, h, w)
        if key not in self._dec:
            self._dec[key] = VaeDecodePlan(self.cfg, self.weights, batch=key[0], frames=num_frames, h=h, w=w,
                                           device=self._device)
        plan = self._dec[key]
        sp = torch.cuda.current_stream().cuda_stream
        zin = z if z.dtype in (F32, BF16) else z.to(F32)
        ops.Layout(zin.contiguous(), plan.z_in, to_tokens=True, halo=True).launch(sp)
        NetPlan.run(plan.ops, sp)
        out = plan.sample.clone()
        return DecoderOutput(sample=out) if return_dict else (out,)

    def forward(self, sample, num_frames: int = 1):
        z = self.encode(sample).latent_dist.mode()
        return self.decode(z, num_frames=num_frames)
