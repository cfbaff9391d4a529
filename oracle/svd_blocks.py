"""ORACLE (test infrastructure, never the product path) — pure-torch CPU restatement of the
diffusers==0.24.0 spatio-temporal UNet blocks that PoseTraj's hot path is wired from.

PARITY UNPINNED: `diffusers` (requirements.txt:4 of the reference) is not installed here and the reference
ships no tests or golden vectors for these blocks.  The semantics below follow
SURVEY.md Appendix A and, where the reference repo carries a patched copy of the same forwards, that copy:
  /root/reference/models/modified_svd.py:50-114   TemporalBasicTransformerBlock.forward
  /root/reference/models/modified_svd.py:118-223  TransformerSpatioTemporalModel.forward
  /root/reference/models/modified_svd.py:225-348  CrossAttn{Up,Down}BlockSpatioTemporal.forward
(minus their `camera_para` lines, which no script of the reference enables).
Module and parameter names equal the diffusers state-dict key tree so a real SVD
checkpoint loads into these modules unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# embeddings (Appendix A.1)
# ----------------------------------------------------------------------------------------------
def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): cat([cos, sin]) in fp32."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    arg = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class Timesteps(nn.Module):
    def __init__(self, num_channels: int):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, t):
        return sinusoidal_embedding(t, self.num_channels)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, out_dim: Optional[int] = None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# ----------------------------------------------------------------------------------------------
# residual blocks (Appendix A.3 - A.5)
# ----------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    """GroupNorm on the 5-D tensor: statistics over (C/32, F, H, W) jointly (Appendix A.4)."""

    def __init__(self, channels: int, temb_channels: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, channels, eps=eps, affine=True)
        self.conv1 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_channels, channels)
        self.norm2 = nn.GroupNorm(32, channels, eps=eps, affine=True)
        self.conv2 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x, temb):  # x [B,C,F,H,W], temb [B,F,T]
        h = self.conv1(F.silu(self.norm1(x)))
        t = self.time_emb_proj(F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
        h = h + t
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class AlphaBlender(nn.Module):
    """merge_strategy="learned_with_images"; image_only_indicator is all zeros on this path."""

    def __init__(self, alpha: float = 0.5):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.tensor([alpha]))

    def get_alpha(self, image_only_indicator, ndims):
        alpha = torch.where(image_only_indicator.bool(),
                            torch.ones(1, 1, device=image_only_indicator.device),
                            torch.sigmoid(self.mix_factor)[..., None])
        if ndims == 5:
            return alpha[:, None, :, None, None]
        return alpha.reshape(-1)[:, None, None]

    def forward(self, x_spatial, x_temporal, image_only_indicator):
        alpha = self.get_alpha(image_only_indicator, x_spatial.ndim).to(x_spatial.dtype)
        return alpha * x_spatial + (1.0 - alpha) * x_temporal


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, eps: float):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, temb_channels, eps)
        self.temporal_res_block = TemporalResnetBlock(out_channels, temb_channels, eps)
        self.time_mixer = AlphaBlender(0.5)

    def forward(self, x, temb, image_only_indicator):
        num_frames = image_only_indicator.shape[-1]
        x = self.spatial_res_block(x, temb)
        bf, c, h, w = x.shape
        b = bf // num_frames
        x_mix = x[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x5 = x[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        temb5 = temb.reshape(b, num_frames, -1)
        x5 = self.temporal_res_block(x5, temb5)
        x5 = self.time_mixer(x_spatial=x_mix, x_temporal=x5, image_only_indicator=image_only_indicator)
        return x5.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


# ----------------------------------------------------------------------------------------------
# attention / feed-forward (Appendix A.7)
# ----------------------------------------------------------------------------------------------
class Attention(nn.Module):
    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int] = None):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, encoder_hidden_states=None):
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        b, s, _ = x.shape
        q = self.to_q(x).view(b, s, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)  # scale = head_dim ** -0.5, no mask, non-causal
        o = o.transpose(1, 2).reshape(b, s, -1)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)  # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h, encoder_hidden_states):
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), encoder_hidden_states) + h
        h = self.ff(self.norm3(h)) + h
        return h


class TemporalBasicTransformerBlock(nn.Module):
    """modified_svd.py:50-114 (time_mix_inner_dim == dim, so is_res is True)."""

    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm_in = nn.LayerNorm(dim, eps=1e-5)
        self.ff_in = FeedForward(dim, dim_out=dim)
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h, num_frames: int, encoder_hidden_states):
        bf, s, c = h.shape
        b = bf // num_frames
        h = h[None, :].reshape(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
        residual = h
        h = self.ff_in(self.norm_in(h)) + residual
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), encoder_hidden_states) + h
        h = self.ff(self.norm3(h)) + h
        return h[None, :].reshape(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(b * num_frames, s, c)


class TransformerSpatioTemporalModel(nn.Module):
    """modified_svd.py:118-223, including the literal (mis-aligned) time_context broadcast (fact 11)."""

    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: int):
        super().__init__()
        inner = heads * dim_head
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.temporal_transformer_blocks = nn.ModuleList(
            [TemporalBasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_proj = Timesteps(in_channels)
        self.time_mixer = AlphaBlender(0.5)
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, encoder_hidden_states, image_only_indicator):
        bf, _, height, width = x.shape
        num_frames = image_only_indicator.shape[-1]
        b = bf // num_frames
        tc = encoder_hidden_states
        tc_first = tc[None, :].reshape(b, num_frames, -1, tc.shape[-1])[:, 0]
        tc = tc_first[None, :].broadcast_to(height * width, b, 1, tc.shape[-1])
        tc = tc.reshape(height * width * b, 1, tc.shape[-1])

        residual = x
        h = self.norm(x)
        inner = h.shape[1]
        h = h.permute(0, 2, 3, 1).reshape(bf, height * width, inner)
        h = self.proj_in(h)

        frames = torch.arange(num_frames, device=x.device).repeat(b, 1).reshape(-1)
        emb = self.time_pos_embed(self.time_proj(frames).to(h.dtype))[:, None, :]

        for block, tblock in zip(self.transformer_blocks, self.temporal_transformer_blocks):
            h = block(h, encoder_hidden_states)
            h_mix = tblock(h + emb, num_frames=num_frames, encoder_hidden_states=tc)
            h = self.time_mixer(x_spatial=h, x_temporal=h_mix, image_only_indicator=image_only_indicator)

        h = self.proj_out(h)
        h = h.reshape(bf, height, width, inner).permute(0, 3, 1, 2).contiguous()
        return h + residual


# ----------------------------------------------------------------------------------------------
# samplers and block stacks (Appendix A.2)
# ----------------------------------------------------------------------------------------------
class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        dtype = x.dtype
        if dtype == torch.bfloat16:  # diffusers upcasts bf16 for nearest interpolation
            x = x.float()
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x.to(dtype))


class CrossAttnDownBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, heads, cross_attention_dim, add_downsample,
                 num_layers=2):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, 1e-6)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(heads, out_channels // heads, out_channels, cross_attention_dim)
            for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, h, temb, encoder_hidden_states, image_only_indicator):
        outs = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            h = resnet(h, temb, image_only_indicator)
            h = attn(h, encoder_hidden_states, image_only_indicator)
            outs += (h,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                h = d(h)
            outs += (h,)
        return h, outs


class DownBlockSpatioTemporal(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, add_downsample, num_layers=2):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, 1e-5)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, h, temb, image_only_indicator):
        outs = ()
        for resnet in self.resnets:
            h = resnet(h, temb, image_only_indicator)
            outs += (h,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                h = d(h)
            outs += (h,)
        return h, outs


class UNetMidBlockSpatioTemporal(nn.Module):
    def __init__(self, channels, temb_channels, heads, cross_attention_dim):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(channels, channels, temb_channels, 1e-5),
                                      SpatioTemporalResBlock(channels, channels, temb_channels, 1e-5)])
        self.attentions = nn.ModuleList(
            [TransformerSpatioTemporalModel(heads, channels // heads, channels, cross_attention_dim)])

    def forward(self, h, temb, encoder_hidden_states, image_only_indicator):
        h = self.resnets[0](h, temb, image_only_indicator)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            h = attn(h, encoder_hidden_states, image_only_indicator)
            h = resnet(h, temb, image_only_indicator)
        return h


class UpBlockSpatioTemporal(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, add_upsample, num_layers=3):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip = in_channels if i == num_layers - 1 else out_channels
            res_in = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(res_in + res_skip, out_channels, temb_channels, 1e-6))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, h, res_hidden_states_tuple, temb, image_only_indicator):
        for resnet in self.resnets:
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            h = resnet(torch.cat([h, skip], dim=1), temb, image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                h = u(h)
        return h


class CrossAttnUpBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, heads, cross_attention_dim,
                 add_upsample, num_layers=3):
        super().__init__()
        resnets, attns = [], []
        for i in range(num_layers):
            res_skip = in_channels if i == num_layers - 1 else out_channels
            res_in = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(res_in + res_skip, out_channels, temb_channels, 1e-6))
            attns.append(TransformerSpatioTemporalModel(heads, out_channels // heads, out_channels, cross_attention_dim))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(attns)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, h, res_hidden_states_tuple, temb, encoder_hidden_states, image_only_indicator):
        for resnet, attn in zip(self.resnets, self.attentions):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            h = resnet(torch.cat([h, skip], dim=1), temb, image_only_indicator)
            h = attn(h, encoder_hidden_states, image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                h = u(h)
        return h
