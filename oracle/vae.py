"""ORACLE (test infrastructure, never the product path) — pure-torch CPU restatement of
`diffusers.AutoencoderKLTemporalDecoder` (diffusers==0.24.0, requirements.txt:4 of the reference), the VAE on either
side of PoseTraj's denoise loop:

  encode   /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:174-195  `_encode_vae_image`
           (`vae.encode(image).latent_dist.mode()`)
  decode   /root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:225-251  `decode_latents`
           (`vae.decode(latents / scaling_factor, num_frames=chunk).sample`, chunked over frames)

PARITY UNPINNED: diffusers is not installed here and the reference ships neither the VAE source nor golden vectors
for it; the block semantics below are the published diffusers 0.24.0 ones (`Encoder`, `TemporalDecoder`,
`MidBlockTemporalDecoder`, `UpBlockTemporalDecoder`, `SpatioTemporalResBlock` with `temb_channels=None`,
`merge_strategy="learned"`, `switch_spatial_to_temporal_mix=True`).  Module and parameter names equal the diffusers
state-dict key tree (`encoder.*`, `decoder.*`, `quant_conv.*`), so a real SVD VAE checkpoint loads unchanged.

Only tests/ and __graft_entry__.smoke() may import this.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlock2D(nn.Module):
    """ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6, output_scale_factor=1)."""

    def __init__(self, in_channels: int, out_channels: int, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    """TemporalResnetBlock(temb_channels=None): GroupNorm statistics over (C/32, F, H, W) of the 5-D tensor."""

    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, channels, eps=eps, affine=True)
        self.conv1 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))
        self.norm2 = nn.GroupNorm(32, channels, eps=eps, affine=True)
        self.conv2 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x):  # [B, C, F, H, W]
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class AlphaBlender(nn.Module):
    """merge_strategy="learned", switch_spatial_to_temporal_mix=True: alpha = 1 - sigmoid(mix_factor)."""

    def __init__(self, alpha: float = 0.0):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.tensor([alpha]))

    def forward(self, x_spatial, x_temporal):
        alpha = 1.0 - torch.sigmoid(self.mix_factor).to(x_spatial.dtype)
        return alpha * x_spatial + (1.0 - alpha) * x_temporal


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, eps=1e-6)
        self.temporal_res_block = TemporalResnetBlock(out_channels, eps=1e-5)
        self.time_mixer = AlphaBlender(0.0)

    def forward(self, x, num_frames: int):
        x = self.spatial_res_block(x)
        bf, c, h, w = x.shape
        b = bf // num_frames
        x5 = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        xt = self.temporal_res_block(x5)
        out = self.time_mixer(x5, xt)
        return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class VaeAttention(nn.Module):
    """Attention(query_dim=C, heads=1, dim_head=C, norm_num_groups=32, eps=1e-6, bias=True, residual_connection=True)
    on a 4-D input: GroupNorm -> q/k/v -> softmax(q k^T / sqrt(C)) v -> to_out -> + input."""

    def __init__(self, channels: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, channels, eps=1e-6, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        n, c, h, w = x.shape
        hs = self.group_norm(x).reshape(n, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(hs), self.to_k(hs), self.to_v(hs)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return x + o.transpose(1, 2).reshape(n, c, h, w)


class _Sampler(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, add_downsample: bool, num_layers: int = 2):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels)
                                      for i in range(num_layers)])
        # Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a stride-2 conv without padding
        self.downsamplers = nn.ModuleList([_Sampler(nn.Conv2d(out_channels, out_channels, 3, stride=2, padding=0))]) \
            if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1)))
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels), ResnetBlock2D(channels, channels)])
        self.attentions = nn.ModuleList([VaeAttention(channels)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self, in_channels=3, latent_channels=4, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block: int = 2):
        super().__init__()
        ch = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        blocks, prev = [], ch[0]
        for i, c in enumerate(ch):
            blocks.append(DownEncoderBlock2D(prev, c, add_downsample=i < len(ch) - 1, num_layers=layers_per_block))
            prev = c
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = UNetMidBlock2D(ch[-1])
        self.conv_norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, channels: int, num_layers: int = 2):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(channels, channels) for _ in range(num_layers)])
        self.attentions = nn.ModuleList([VaeAttention(channels)])

    def forward(self, x, num_frames):
        x = self.resnets[0](x, num_frames)
        for resnet, attn in zip(self.resnets[1:], self.attentions):
            x = attn(x)
            x = resnet(x, num_frames)
        return x


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels, out_channels, add_upsample: bool, num_layers: int = 3):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([_Sampler(nn.Conv2d(out_channels, out_channels, 3, padding=1))]) \
            if add_upsample else None

    def forward(self, x, num_frames):
        for r in self.resnets:
            x = r(x, num_frames)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class TemporalDecoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block: int = 2):
        super().__init__()
        ch = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, ch[-1], 3, padding=1)
        self.mid_block = MidBlockTemporalDecoder(ch[-1], num_layers=layers_per_block)
        rev = list(reversed(ch))
        blocks, prev = [], rev[0]
        for i, c in enumerate(rev):
            blocks.append(UpBlockTemporalDecoder(prev, c, add_upsample=i < len(rev) - 1, num_layers=layers_per_block + 1))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, z, num_frames: int):
        x = self.conv_in(z)
        x = self.mid_block(x, num_frames)
        for b in self.up_blocks:
            x = b(x, num_frames)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        bf, c, h, w = x.shape
        x = x.reshape(bf // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x = self.time_conv_out(x)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class AutoencoderKLTemporalDecoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block: int = 2, latent_channels: int = 4, scaling_factor: float = 0.18215,
                 force_upcast: bool = True):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.decoder = TemporalDecoder(latent_channels, out_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      latent_channels=latent_channels, scaling_factor=scaling_factor,
                                      force_upcast=force_upcast)

    def encode_mode(self, x):
        """`vae.encode(x).latent_dist.mode()`: the mean half of quant_conv(encoder(x))."""
        moments = self.quant_conv(self.encoder(x))
        return moments[:, : moments.shape[1] // 2]

    def decode(self, z, num_frames: int):
        """`vae.decode(z, num_frames=num_frames).sample` (image_only_indicator is all zeros and unused by the
        "learned" blender)."""
        return self.decoder(z, num_frames)


def build_vae(seed: int = 0, block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2,
              randomize_mix: bool = True) -> AutoencoderKLTemporalDecoder:
    """Random-init VAE under a fixed seed (torch default inits; mix factors drawn ~N(0,1) so that the blend is not
    the symmetric 0.5 everywhere)."""
    torch.manual_seed(seed)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=block_out_channels, layers_per_block=layers_per_block)
    if randomize_mix:
        g = torch.Generator().manual_seed(seed + 1)
        for n, p in vae.named_parameters():
            if n.endswith("mix_factor"):
                p.data.copy_(torch.randn(1, generator=g))
    return vae.eval()
