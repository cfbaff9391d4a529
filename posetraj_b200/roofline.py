"""Algorithmic FLOP / byte counts of one denoise step — the numerators of every roofline
fraction bench.py reports.

FLOPs = 2 x multiply-accumulates of the dense contractions only (convs, linears, attention cores); biases, norms,
activations, blends, softmax and embeddings-of-scalars are excluded.  The walk mirrors the architecture of
  /root/reference/models/unet_spatio_temporal_condition_controlnet.py:126-245  (UNet __init__)
  /root/reference/models/controlnet_sdv.py:299-391                             (ControlNet __init__)
`essential=True` drops work the reference spends on the degenerate 1-token cross-attention:
the dead to_q GEMMs and the per-token to_out GEMMs (their result is a per-batch constant vector).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Tuple

from .config import SVDConfig, up_block_plan


class _Counter:
    def __init__(self, cfg: SVDConfig, B: int, F: int, h: int, w: int, essential: bool):
        self.cfg, self.B, self.F, self.h, self.w, self.essential = cfg, B, F, h, w, essential
        self.BF = B * F
        self.cat: Dict[str, float] = defaultdict(float)

    def add(self, name: str, flops: float) -> None:
        self.cat[name] += flops

    def conv(self, name, k, cin, cout, ho, wo):
        self.add(name, 2.0 * k * k * cin * cout * ho * wo * self.BF)

    def resblock(self, cin, cout, h, w):
        temb = self.cfg.temb_dim
        self.conv("conv3x3", 3, cin, cout, h, w)
        self.conv("conv3x3", 3, cout, cout, h, w)
        if cin != cout:
            self.conv("shortcut", 1, cin, cout, h, w)
        self.add("temporal_conv", 2 * (2.0 * 3 * cout * cout * h * w * self.BF))
        self.add("time_emb_proj", 2 * (2.0 * temb * cout * self.BF))

    def ff(self, d, T):
        self.add("geglu_in", 2.0 * d * 8 * d * T)
        self.add("ff_out", 2.0 * 4 * d * d * T)

    def transformer(self, d, h, w):
        N = h * w
        T = self.BF * N
        x = self.cfg.cross_attention_dim
        self.add("proj_in_out", 2 * (2.0 * d * d * T))
        # spatial block
        self.add("spatial_qkv", 3 * 2.0 * d * d * T)
        self.add("attn_out_proj", 2.0 * d * d * T)
        self.add("spatial_attn_core", 4.0 * N * d * T)
        if not self.essential:
            self.add("xattn_dead_q", 2.0 * d * d * T)
            self.add("xattn_degenerate_out", 2.0 * d * d * T)
            self.add("xattn_kv", 2 * 2.0 * x * d * self.BF)
        self.ff(d, T)
        # temporal block
        self.ff(d, T)  # ff_in
        self.add("temporal_qkv", 3 * 2.0 * d * d * T)
        self.add("attn_out_proj", 2.0 * d * d * T)
        self.add("temporal_attn_core", 4.0 * self.F * d * T)
        if not self.essential:
            self.add("xattn_dead_q", 2.0 * d * d * T)
            self.add("xattn_degenerate_out", 2.0 * d * d * T)
            self.add("xattn_kv", 2 * 2.0 * x * d * self.B * N)
        self.ff(d, T)
        self.add("time_pos_embed", 2.0 * (4 * d * d + 4 * d * d) * self.BF)

    def time_embeddings(self):
        c = self.cfg
        t = c.temb_dim
        self.add("time_embeddings", 2.0 * (c.block_out_channels[0] * t + t * t +
                                          c.projection_class_embeddings_input_dim * t + t * t) * self.B)

    def encoder(self):
        c = self.cfg
        ch = c.block_out_channels
        n = len(ch)
        self.time_embeddings()
        self.conv("conv_in_out", 3, c.in_channels, ch[0], self.h, self.w)
        prev = ch[0]
        for i in range(n):
            h, w = self.h >> i, self.w >> i
            for _ in range(c.layers_per_block):
                self.resblock(prev, ch[i], h, w)
                prev = ch[i]
                if i < n - 1:
                    self.transformer(ch[i], h, w)
            if i < n - 1:
                self.conv("downsample", 3, ch[i], ch[i], h >> 1, w >> 1)
        h, w = self.h >> (n - 1), self.w >> (n - 1)
        self.resblock(ch[-1], ch[-1], h, w)
        self.transformer(ch[-1], h, w)
        self.resblock(ch[-1], ch[-1], h, w)

    def decoder(self):
        c = self.cfg
        ch = c.block_out_channels
        n = len(ch)
        for i, (out_c, layers, has_attn, add_up) in enumerate(up_block_plan(c)):
            lvl = n - 1 - i
            h, w = self.h >> lvl, self.w >> lvl
            for (res_in, skip_c) in layers:
                self.resblock(res_in + skip_c, out_c, h, w)
                if has_attn:
                    self.transformer(out_c, h, w)
            if add_up:
                self.conv("upsample", 3, out_c, out_c, 2 * h, 2 * w)
        self.conv("conv_in_out", 3, ch[0], c.out_channels, self.h, self.w)

    def controlnet_extras(self, cam: bool, bbox: bool):
        c = self.cfg
        ch = c.block_out_channels
        n = len(ch)
        # 12 + 1 zero convs at the skip shapes
        self.conv("zero_convs", 1, ch[0], ch[0], self.h, self.w)
        for i in range(n):
            h, w = self.h >> i, self.w >> i
            for _ in range(c.layers_per_block):
                self.conv("zero_convs", 1, ch[i], ch[i], h, w)
            if i < n - 1:
                self.conv("zero_convs", 1, ch[i], ch[i], h >> 1, w >> 1)
        self.conv("zero_convs", 1, ch[-1], ch[-1], self.h >> (n - 1), self.w >> (n - 1))
        # conditioning embedding at pixel resolution (controlnet_sdv.py:84-109)
        ce = c.conditioning_embedding_out_channels
        H, W = self.h * 8, self.w * 8
        towers = 2 if bbox else 1
        for _ in range(towers):
            self.conv("cond_embed", 3, c.conditioning_channels, ce[0], H, W)
            hh, ww = H, W
            for i in range(len(ce) - 1):
                self.conv("cond_embed", 3, ce[i], ce[i], hh, ww)
                hh, ww = hh // 2, ww // 2
                self.conv("cond_embed", 3, ce[i], ce[i + 1], hh, ww)
            self.conv("cond_embed", 3, ce[-1], ch[0], hh, ww)
        if cam:
            self.add("cc_projection", 2.0 * (ce[-1] + 12) * ce[-1] * self.h * self.w * self.BF)


def step_flops(cfg: SVDConfig = SVDConfig(), *, batch: int = 2, frames: int = 14, h: int = 40, w: int = 72,
               cam: bool = False, bbox: bool = False, essential: bool = False) -> Tuple[float, Dict[str, float]]:
    """FLOPs of one denoise step (ControlNet + UNet) and their per-category split.

    The conditioning embedding is counted like the reference runs it (every step); the product hoists it out of the
    loop (it is step-invariant), which is why bench.py uses the `essential` count minus nothing else — conservative."""
    u = _Counter(cfg, batch, frames, h, w, essential)
    u.encoder()
    u.decoder()
    cn = _Counter(cfg, batch, frames, h, w, essential)
    cn.encoder()
    cn.controlnet_extras(cam, bbox)
    cats: Dict[str, float] = defaultdict(float)
    for src in (u.cat, cn.cat):
        for k, v in src.items():
            cats[k] += v
    cats["_unet"] = sum(u.cat.values())
    cats["_controlnet"] = sum(cn.cat.values())
    return cats["_unet"] + cats["_controlnet"], dict(cats)


def glue_bytes(cfg: SVDConfig = SVDConfig(), *, batch: int = 2, frames: int = 14, h: int = 40, w: int = 72) -> Dict[str, float]:
    """Read-once / write-once bf16 bytes of the fused CFG+Euler kernel."""
    lat = frames * cfg.out_channels * h * w
    return {"cfg_euler": lat * (2 * 2 + 4 + 4)}
