"""ORACLE (test infrastructure, never the product path) — CPU restatement of PoseTraj's model wiring.

PARITY UNPINNED for the numerics of the leaf blocks (see oracle/svd_blocks.py).  The WIRING restated here is PINNED:
tests/golden/wiring_golden.safetensors holds outputs of the reference's own model files executed unmodified with
`diffusers` shimmed (tests/golden/gen_wiring_golden.py), and tests/test_wiring_golden_cpu.py asserts this module
reproduces them (plain / camera / bbox variants, 13 residuals + noise prediction, <= 1e-5).  It follows the reference
files line by line:
  /root/reference/models/controlnet_sdv.py:61-116      ControlNetConditioningEmbeddingSVD
  /root/reference/models/controlnet_sdv.py:299-391     ControlNetSDVModel.__init__
  /root/reference/models/controlnet_sdv.py:516-650     ControlNetSDVModel.forward
  /root/reference/models/controlnet_sdv_cam_infer.py:84,96-122   camera branch (cc_projection)
  /root/reference/models/controlnet_sdv_bbox.py:95-138           bbox tower (shares conv_out, :134)
  /root/reference/models/unet_spatio_temporal_condition_controlnet.py:126-245   UNet __init__
  /root/reference/models/unet_spatio_temporal_condition_controlnet.py:386-504   UNet forward, including the
      in-loop residual accumulation (:451-459, effective multipliers [4,4,4,4,3,3,3,2,2,2,1,1]).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .svd_blocks import (CrossAttnDownBlockSpatioTemporal, CrossAttnUpBlockSpatioTemporal, DownBlockSpatioTemporal,
                         TimestepEmbedding, Timesteps, UNetMidBlockSpatioTemporal, UpBlockSpatioTemporal)

SVD_DEFAULTS = dict(
    in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    num_attention_heads=(5, 10, 20, 20), num_frames=14,
)


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


class _EncoderTrunk(nn.Module):
    """conv_in + time/aug embeddings + down blocks + mid block: shared by the UNet and the ControlNet."""

    def __init__(self, in_channels, block_out_channels, addition_time_embed_dim,
                 projection_class_embeddings_input_dim, layers_per_block, cross_attention_dim, num_attention_heads):
        super().__init__()
        c0 = block_out_channels[0]
        temb = c0 * 4
        self.conv_in = nn.Conv2d(in_channels, c0, 3, padding=1)
        self.time_proj = Timesteps(c0)
        self.time_embedding = TimestepEmbedding(c0, temb)
        self.add_time_proj = Timesteps(addition_time_embed_dim)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb)
        self.temb_dim = temb

    def _build_down(self, block_out_channels, layers_per_block, cross_attention_dim, num_attention_heads):
        blocks = nn.ModuleList()
        out_c = block_out_channels[0]
        n = len(block_out_channels)
        for i in range(n):
            in_c, out_c = out_c, block_out_channels[i]
            final = i == n - 1
            if i < n - 1:
                blocks.append(CrossAttnDownBlockSpatioTemporal(in_c, out_c, self.temb_dim, num_attention_heads[i],
                                                               cross_attention_dim, add_downsample=not final,
                                                               num_layers=layers_per_block))
            else:
                blocks.append(DownBlockSpatioTemporal(in_c, out_c, self.temb_dim, add_downsample=not final,
                                                      num_layers=layers_per_block))
        return blocks

    def embed(self, sample, timestep, added_time_ids):
        """time + aug embedding, repeated per frame (controlnet_sdv.py:550-590)."""
        b, f = sample.shape[:2]
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float64 if isinstance(t, float) else torch.int64, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t.expand(b)
        emb = self.time_embedding(self.time_proj(t).to(sample.dtype))
        te = self.add_time_proj(added_time_ids.flatten()).reshape(b, -1).to(emb.dtype)
        emb = emb + self.add_embedding(te)
        return emb.repeat_interleave(f, dim=0)


class ControlNetConditioningEmbeddingSVD(nn.Module):
    """controlnet_sdv.py:61-116; `cam=True` adds cc_projection (controlnet_sdv_cam_infer.py:84,109-119);
    `bbox=True` adds the second tower of controlnet_sdv_bbox.py:95-138."""

    def __init__(self, out_channels: int, cond_channels: int = 3, chans: Sequence[int] = (16, 32, 96, 256),
                 cam: bool = False, bbox: bool = False):
        super().__init__()
        self.conv_in = nn.Conv2d(cond_channels, chans[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        if cam:
            self.cc_projection = nn.Linear(chans[-1] + 12, chans[-1])
        for i in range(len(chans) - 1):
            self.blocks.append(nn.Conv2d(chans[i], chans[i], 3, padding=1))
            self.blocks.append(nn.Conv2d(chans[i], chans[i + 1], 3, padding=1, stride=2))
        self.conv_out = zero_module(nn.Conv2d(chans[-1], out_channels, 3, padding=1))
        self.bbox = bbox
        if bbox:
            self.conv_in_2 = nn.Conv2d(cond_channels, chans[0], 3, padding=1)
            self.blocks_2 = nn.ModuleList()
            for i in range(len(chans) - 1):
                self.blocks_2.append(nn.Conv2d(chans[i], chans[i], 3, padding=1))
                self.blocks_2.append(nn.Conv2d(chans[i], chans[i + 1], 3, padding=1, stride=2))
            self.conv_out_2 = zero_module(nn.Conv2d(chans[-1], out_channels, 3, padding=1))  # dead in the reference

    @staticmethod
    def _tower(x, conv_in, blocks):
        e = F.silu(conv_in(x))
        for blk in blocks:
            e = F.silu(blk(e))
        return e

    def forward(self, conditioning, camera_RT=None, conditioning_bbox=None):
        b, f, c, h, w = conditioning.shape
        e = self._tower(conditioning.reshape(b * f, c, h, w), self.conv_in, self.blocks)
        if camera_RT is not None:
            cam = camera_RT.reshape(b * f, -1)[:, :, None, None].repeat(1, 1, e.shape[2], e.shape[3])
            e = torch.cat((e, cam), dim=1).permute(0, 2, 3, 1).contiguous()
            e = self.cc_projection(e).permute(0, 3, 1, 2).contiguous()
        e = self.conv_out(e)
        if self.bbox and conditioning_bbox is not None:
            e2 = self._tower(conditioning_bbox.reshape(b * f, c, h, w), self.conv_in_2, self.blocks_2)
            e = e + self.conv_out(e2)  # the reference projects tower 2 with the SHARED conv_out (bbox.py:134)
        return e


class ControlNetSDVModel(_EncoderTrunk):
    def __init__(self, in_channels=8, block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
                 projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
                 num_attention_heads=(5, 10, 20, 20), num_frames=14, conditioning_channels=3,
                 conditioning_embedding_out_channels=(16, 32, 96, 256), cam=False, bbox=False, out_channels=4):
        super().__init__(in_channels, block_out_channels, addition_time_embed_dim,
                         projection_class_embeddings_input_dim, layers_per_block, cross_attention_dim,
                         num_attention_heads)
        # construction order follows controlnet_sdv.py:299-391 (matters for seeded init)
        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList()
        self.controlnet_cond_embedding = ControlNetConditioningEmbeddingSVD(
            block_out_channels[0], conditioning_channels, conditioning_embedding_out_channels, cam=cam, bbox=bbox)
        c = block_out_channels[0]
        self.controlnet_down_blocks.append(zero_module(nn.Conv2d(c, c, 1)))
        downs = self._build_down(block_out_channels, layers_per_block, cross_attention_dim, num_attention_heads)
        n = len(block_out_channels)
        for i, blk in enumerate(downs):
            self.down_blocks.append(blk)
            c = block_out_channels[i]
            for _ in range(layers_per_block):
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(c, c, 1)))
            if i != n - 1:
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(c, c, 1)))
        self.controlnet_mid_block = zero_module(nn.Conv2d(block_out_channels[-1], block_out_channels[-1], 1))
        self.mid_block = UNetMidBlockSpatioTemporal(block_out_channels[-1], self.temb_dim, num_attention_heads[-1],
                                                    cross_attention_dim)

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, controlnet_cond=None,
                camera_cond=None, controlnet_bbox=None, image_only_indicator=None, return_dict=True,
                guess_mode=False, conditioning_scale=1.0):
        b, f = sample.shape[:2]
        emb = self.embed(sample, timestep, added_time_ids)
        x = sample.flatten(0, 1)
        ehs = encoder_hidden_states.repeat_interleave(f, dim=0)
        x = self.conv_in(x)
        if controlnet_cond is not None:
            x = x + self.controlnet_cond_embedding(controlnet_cond, camera_cond, controlnet_bbox)
        ioi = torch.zeros(b, f, dtype=x.dtype, device=x.device)
        skips = (x,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                x, res = blk(x, emb, ehs, ioi)
            else:
                x, res = blk(x, emb, ioi)
            skips += res
        x = self.mid_block(x, emb, ehs, ioi)
        down = [conv(s) * conditioning_scale for s, conv in zip(skips, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(x) * conditioning_scale
        return down, mid


class UNetSpatioTemporalConditionControlNetModel(_EncoderTrunk):
    def __init__(self, in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                 addition_time_embed_dim=256, projection_class_embeddings_input_dim=768, layers_per_block=2,
                 cross_attention_dim=1024, num_attention_heads=(5, 10, 20, 20), num_frames=14):
        super().__init__(in_channels, block_out_channels, addition_time_embed_dim,
                         projection_class_embeddings_input_dim, layers_per_block, cross_attention_dim,
                         num_attention_heads)
        self.down_blocks = self._build_down(block_out_channels, layers_per_block, cross_attention_dim,
                                            num_attention_heads)
        self.up_blocks = nn.ModuleList()
        self.mid_block = UNetMidBlockSpatioTemporal(block_out_channels[-1], self.temb_dim, num_attention_heads[-1],
                                                    cross_attention_dim)
        rev_c = list(reversed(block_out_channels))
        rev_h = list(reversed(num_attention_heads))
        n = len(block_out_channels)
        out_c = rev_c[0]
        for i in range(n):
            prev, out_c = out_c, rev_c[i]
            in_c = rev_c[min(i + 1, n - 1)]
            final = i == n - 1
            if i == 0:
                self.up_blocks.append(UpBlockSpatioTemporal(in_c, prev, out_c, self.temb_dim, add_upsample=not final,
                                                            num_layers=layers_per_block + 1))
            else:
                self.up_blocks.append(CrossAttnUpBlockSpatioTemporal(in_c, prev, out_c, self.temb_dim, rev_h[i],
                                                                     cross_attention_dim, add_upsample=not final,
                                                                     num_layers=layers_per_block + 1))
        self.conv_norm_out = nn.GroupNorm(32, block_out_channels[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True, added_time_ids=None):
        b, f = sample.shape[:2]
        emb = self.embed(sample, timestep, added_time_ids)
        x = sample.flatten(0, 1)
        ehs = encoder_hidden_states.repeat_interleave(f, dim=0)
        x = self.conv_in(x)
        ioi = torch.zeros(b, f, dtype=x.dtype, device=x.device)
        skips = (x,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                x, res = blk(x, emb, ehs, ioi)
            else:
                x, res = blk(x, emb, ioi)
            skips += res
            # the add sits INSIDE the block loop in the reference (:451-459): earlier skips get it repeatedly
            if down_block_additional_residuals is not None:
                skips = tuple(s + r for s, r in zip(skips, down_block_additional_residuals))
        x = self.mid_block(x, emb, ehs, ioi)
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual
        for blk in self.up_blocks:
            k = len(blk.resnets)
            res, skips = skips[-k:], skips[:-k]
            if blk.has_cross_attention:
                x = blk(x, res, emb, ehs, ioi)
            else:
                x = blk(x, res, emb, ioi)
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        return x.reshape(b, f, *x.shape[1:])


def build_models(seed: int = 0, cam: bool = False, bbox: bool = False, randomize_zero_convs: bool = False,
                 **cfg) -> Tuple[UNetSpatioTemporalConditionControlNetModel, ControlNetSDVModel]:
    """Seeded random-init pair: W0 = faithful init (zero convs are zero);
    `randomize_zero_convs` gives W1: the 13 zero-convs and cond_embedding.conv_out re-drawn N(0, 1/fan_in), seed+1."""
    kw = dict(SVD_DEFAULTS)
    kw.update(cfg)
    torch.manual_seed(seed)
    unet = UNetSpatioTemporalConditionControlNetModel(**kw).eval()
    ckw = {k: v for k, v in kw.items() if k != "out_channels"}
    cnet = ControlNetSDVModel(cam=cam, bbox=bbox, **ckw).eval()
    if cam:
        # training-time init of the camera projection: identity on the feature block, zero bias
        # (scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1002-1004); the 12 camera columns keep the default init
        with torch.no_grad():
            pj = cnet.controlnet_cond_embedding.cc_projection
            nn.init.eye_(pj.weight[:, : pj.out_features])
            nn.init.zeros_(pj.bias)
    if randomize_zero_convs:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            convs = list(cnet.controlnet_down_blocks) + [cnet.controlnet_mid_block,
                                                         cnet.controlnet_cond_embedding.conv_out]
            for conv in convs:
                fan_in = conv.weight[0].numel()
                conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / fan_in ** 0.5)
                conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.02)
    for p in list(unet.parameters()) + list(cnet.parameters()):
        p.requires_grad_(False)
    return unet, cnet
