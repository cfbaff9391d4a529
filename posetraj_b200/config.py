"""Model configuration and the diffusers state-dict key tree of the two networks on the hot path.

`SVDConfig` mirrors the constructor kwargs of the reference classes
(/root/reference/models/controlnet_sdv.py:238-262, models/unet_spatio_temporal_condition_controlnet.py:68-100)
with the values of the `stabilityai/stable-video-diffusion-img2vid` checkpoint (note the heads:
the class default (5,10,10,20) differs from the checkpoint's (5,10,20,20)).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple


@dataclass(frozen=True)
class SVDConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 768
    num_frames: int = 14
    sample_size: int = 96
    transformer_layers_per_block: int = 1
    conditioning_channels: int = 3
    conditioning_embedding_out_channels: Tuple[int, ...] = (16, 32, 96, 256)
    time_cond_proj_dim: None = None

    def __post_init__(self):
        if len(self.block_out_channels) != len(self.num_attention_heads):
            raise ValueError("Must provide the same number of `num_attention_heads` as `block_out_channels`.")
        for c, h in zip(self.block_out_channels, self.num_attention_heads):
            if c != 64 * h:
                raise ValueError("posetraj_b200 kernels need head_dim == 64 at every level (C == 64*heads)")

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * 4


def _resblock(shapes: Dict, p: str, cin: int, cout: int, temb: int) -> None:
    s, t = p + "spatial_res_block.", p + "temporal_res_block."
    shapes[s + "norm1.weight"] = (cin,); shapes[s + "norm1.bias"] = (cin,)
    shapes[s + "conv1.weight"] = (cout, cin, 3, 3); shapes[s + "conv1.bias"] = (cout,)
    shapes[s + "time_emb_proj.weight"] = (cout, temb); shapes[s + "time_emb_proj.bias"] = (cout,)
    shapes[s + "norm2.weight"] = (cout,); shapes[s + "norm2.bias"] = (cout,)
    shapes[s + "conv2.weight"] = (cout, cout, 3, 3); shapes[s + "conv2.bias"] = (cout,)
    if cin != cout:
        shapes[s + "conv_shortcut.weight"] = (cout, cin, 1, 1); shapes[s + "conv_shortcut.bias"] = (cout,)
    shapes[t + "norm1.weight"] = (cout,); shapes[t + "norm1.bias"] = (cout,)
    shapes[t + "conv1.weight"] = (cout, cout, 3, 1, 1); shapes[t + "conv1.bias"] = (cout,)
    shapes[t + "time_emb_proj.weight"] = (cout, temb); shapes[t + "time_emb_proj.bias"] = (cout,)
    shapes[t + "norm2.weight"] = (cout,); shapes[t + "norm2.bias"] = (cout,)
    shapes[t + "conv2.weight"] = (cout, cout, 3, 1, 1); shapes[t + "conv2.bias"] = (cout,)
    shapes[p + "time_mixer.mix_factor"] = (1,)


def _attn(shapes: Dict, p: str, c: int, kv: int) -> None:
    shapes[p + "to_q.weight"] = (c, c)
    shapes[p + "to_k.weight"] = (c, kv)
    shapes[p + "to_v.weight"] = (c, kv)
    shapes[p + "to_out.0.weight"] = (c, c); shapes[p + "to_out.0.bias"] = (c,)


def _ff(shapes: Dict, p: str, c: int) -> None:
    shapes[p + "net.0.proj.weight"] = (8 * c, c); shapes[p + "net.0.proj.bias"] = (8 * c,)
    shapes[p + "net.2.weight"] = (c, 4 * c); shapes[p + "net.2.bias"] = (c,)


def _transformer(shapes: Dict, p: str, c: int, xdim: int) -> None:
    shapes[p + "norm.weight"] = (c,); shapes[p + "norm.bias"] = (c,)
    shapes[p + "proj_in.weight"] = (c, c); shapes[p + "proj_in.bias"] = (c,)
    b = p + "transformer_blocks.0."
    for n in ("norm1", "norm2", "norm3"):
        shapes[b + n + ".weight"] = (c,); shapes[b + n + ".bias"] = (c,)
    _attn(shapes, b + "attn1.", c, c)
    _attn(shapes, b + "attn2.", c, xdim)
    _ff(shapes, b + "ff.", c)
    t = p + "temporal_transformer_blocks.0."
    for n in ("norm_in", "norm1", "norm2", "norm3"):
        shapes[t + n + ".weight"] = (c,); shapes[t + n + ".bias"] = (c,)
    _ff(shapes, t + "ff_in.", c)
    _attn(shapes, t + "attn1.", c, c)
    _attn(shapes, t + "attn2.", c, xdim)
    _ff(shapes, t + "ff.", c)
    shapes[p + "time_pos_embed.linear_1.weight"] = (4 * c, c); shapes[p + "time_pos_embed.linear_1.bias"] = (4 * c,)
    shapes[p + "time_pos_embed.linear_2.weight"] = (c, 4 * c); shapes[p + "time_pos_embed.linear_2.bias"] = (c,)
    shapes[p + "time_mixer.mix_factor"] = (1,)
    shapes[p + "proj_out.weight"] = (c, c); shapes[p + "proj_out.bias"] = (c,)


def _trunk(shapes: Dict, cfg: SVDConfig) -> None:
    ch, temb, xdim = cfg.block_out_channels, cfg.temb_dim, cfg.cross_attention_dim
    shapes["conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3); shapes["conv_in.bias"] = (ch[0],)
    for name, cin in (("time_embedding", ch[0]), ("add_embedding", cfg.projection_class_embeddings_input_dim)):
        shapes[f"{name}.linear_1.weight"] = (temb, cin); shapes[f"{name}.linear_1.bias"] = (temb,)
        shapes[f"{name}.linear_2.weight"] = (temb, temb); shapes[f"{name}.linear_2.bias"] = (temb,)
    n = len(ch)
    out_c = ch[0]
    for i in range(n):
        in_c, out_c = out_c, ch[i]
        for j in range(cfg.layers_per_block):
            _resblock(shapes, f"down_blocks.{i}.resnets.{j}.", in_c if j == 0 else out_c, out_c, temb)
            if i < n - 1:
                _transformer(shapes, f"down_blocks.{i}.attentions.{j}.", out_c, xdim)
        if i < n - 1:
            shapes[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (out_c, out_c, 3, 3)
            shapes[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (out_c,)
    _resblock(shapes, "mid_block.resnets.0.", ch[-1], ch[-1], temb)
    _transformer(shapes, "mid_block.attentions.0.", ch[-1], xdim)
    _resblock(shapes, "mid_block.resnets.1.", ch[-1], ch[-1], temb)


def up_block_plan(cfg: SVDConfig):
    """[(level_channels_out, [(resnet_in_channels, skip_channels)], has_attention, add_upsample)] for the 4 up blocks
    (models/unet_spatio_temporal_condition_controlnet.py:197-234 + diffusers block factory)."""
    ch = cfg.block_out_channels
    rev = list(reversed(ch))
    n = len(ch)
    plan = []
    out_c = rev[0]
    for i in range(n):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, n - 1)]
        layers = []
        for j in range(cfg.layers_per_block + 1):
            skip = in_c if j == cfg.layers_per_block else out_c
            res_in = prev if j == 0 else out_c
            layers.append((res_in, skip))
        plan.append((out_c, layers, i > 0, i < n - 1))
    return plan


def unet_param_shapes(cfg: SVDConfig) -> Dict[str, tuple]:
    shapes: Dict[str, tuple] = {}
    _trunk(shapes, cfg)
    for i, (out_c, layers, has_attn, add_up) in enumerate(up_block_plan(cfg)):
        for j, (res_in, skip) in enumerate(layers):
            _resblock(shapes, f"up_blocks.{i}.resnets.{j}.", res_in + skip, out_c, cfg.temb_dim)
            if has_attn:
                _transformer(shapes, f"up_blocks.{i}.attentions.{j}.", out_c, cfg.cross_attention_dim)
        if add_up:
            shapes[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (out_c, out_c, 3, 3)
            shapes[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (out_c,)
    c0 = cfg.block_out_channels[0]
    shapes["conv_norm_out.weight"] = (c0,); shapes["conv_norm_out.bias"] = (c0,)
    shapes["conv_out.weight"] = (cfg.out_channels, c0, 3, 3); shapes["conv_out.bias"] = (cfg.out_channels,)
    return shapes


def controlnet_param_shapes(cfg: SVDConfig, cam: bool = False, bbox: bool = False) -> Dict[str, tuple]:
    shapes: Dict[str, tuple] = {}
    _trunk(shapes, cfg)
    ch = cfg.block_out_channels
    ce = cfg.conditioning_embedding_out_channels
    p = "controlnet_cond_embedding."
    towers = [("conv_in", "blocks", "conv_out")] + ([("conv_in_2", "blocks_2", "conv_out_2")] if bbox else [])
    for cin_name, blocks_name, cout_name in towers:
        shapes[p + cin_name + ".weight"] = (ce[0], cfg.conditioning_channels, 3, 3); shapes[p + cin_name + ".bias"] = (ce[0],)
        for i in range(len(ce) - 1):
            shapes[p + f"{blocks_name}.{2 * i}.weight"] = (ce[i], ce[i], 3, 3); shapes[p + f"{blocks_name}.{2 * i}.bias"] = (ce[i],)
            shapes[p + f"{blocks_name}.{2 * i + 1}.weight"] = (ce[i + 1], ce[i], 3, 3)
            shapes[p + f"{blocks_name}.{2 * i + 1}.bias"] = (ce[i + 1],)
        shapes[p + cout_name + ".weight"] = (ch[0], ce[-1], 3, 3); shapes[p + cout_name + ".bias"] = (ch[0],)
    if cam:
        shapes[p + "cc_projection.weight"] = (ce[-1], ce[-1] + 12); shapes[p + "cc_projection.bias"] = (ce[-1],)
    idx = 0
    shapes[f"controlnet_down_blocks.{idx}.weight"] = (ch[0], ch[0], 1, 1); shapes[f"controlnet_down_blocks.{idx}.bias"] = (ch[0],)
    idx += 1
    n = len(ch)
    for i in range(n):
        for _ in range(cfg.layers_per_block + (1 if i < n - 1 else 0)):
            shapes[f"controlnet_down_blocks.{idx}.weight"] = (ch[i], ch[i], 1, 1)
            shapes[f"controlnet_down_blocks.{idx}.bias"] = (ch[i],)
            idx += 1
    shapes["controlnet_mid_block.weight"] = (ch[-1], ch[-1], 1, 1); shapes["controlnet_mid_block.bias"] = (ch[-1],)
    return shapes


RESIDUAL_MULTIPLIERS_DOC = """The UNet adds the ControlNet residuals INSIDE its down-block loop
(models/unet_spatio_temporal_condition_controlnet.py:451-459), so skip i receives its residual once per remaining
block: with 4 blocks and (3,3,3,2) skips per block (+conv_in) the effective multipliers are
[4,4,4,4,3,3,3,2,2,2,1,1]."""


def residual_multipliers(cfg: SVDConfig) -> list:
    n = len(cfg.block_out_channels)
    counts = [1 + cfg.layers_per_block + 1] + [cfg.layers_per_block + (1 if i < n - 1 else 0) for i in range(1, n)]
    mult = []
    for blk, cnt in enumerate(counts):
        mult += [n - blk] * cnt
    return mult
