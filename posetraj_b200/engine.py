"""Lowering of the two networks of PoseTraj's denoise step onto the sm_100a kernel library.

`NetPlan` walks the architecture once for a fixed (batch, frames, height, width) and emits flat lists of
pre-built launch descriptors (posetraj_b200.ops): running a forward is replaying a list, one ctypes call per
kernel, no tensor math on the host.  The wiring follows the reference files
  /root/reference/models/controlnet_sdv.py:516-650 (ControlNetSDVModel.forward)
  /root/reference/models/unet_spatio_temporal_condition_controlnet.py:386-504 (UNet forward)
  /root/reference/models/modified_svd.py:50-348 (diffusers block forwards)
with these algebraic rewrites, each exact in real arithmetic:
  * activations are token-major NHWC bf16; `[B*F, C, H, W] <-> [B, C, F, H, W]` reshapes of the reference are
    the same rows here, so they disappear;
  * 1-token cross-attention == a constant vector `to_out(to_v(e_b))` per batch row, pre-computed when the image
    embedding changes and added in the epilogue of the preceding GEMM.  The temporal block receives the vector of
    batch `(b*HW + s) mod B` — the reference's (mis-aligned) broadcast, reproduced on purpose;
  * AlphaBlender(x, temporal(x)) with temporal(x) = x + f(x) is `x + (1-alpha) f(x)`: an epilogue scale;
  * `torch.cat([h, skip], 1)` is never materialised: GroupNorm and the 1x1 shortcut read both sources;
  * the ControlNet residuals enter the UNet as `skip_i = h_i + m_i r_i` with m = [4,4,4,4,3,3,3,2,2,2,1,1]
    (the reference's in-loop accumulation), fused into the epilogue of the GEMM that produces h_i;
  * time_emb_proj(silu(emb)) of all resnets of a network is one batched GEMV per step.
"""
from __future__ import annotations

import math
import os
from collections import defaultdict
from typing import Dict, List, Optional

import torch

from . import ops
from .config import SVDConfig, residual_multipliers, up_block_plan

BF16 = torch.bfloat16
F32 = torch.float32
# experiment knob: PT_FUSED_MLP=0 keeps the GEGLU GEMM + output GEMM pair at every width (A/B runs)
FUSED_MLP = os.environ.get("PT_FUSED_MLP", "1") != "0"


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class WeightStore:
    """Device-side, kernel-ready views of a diffusers-format state dict (SURVEY.md Appendix E key tree)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device):
        self.sd = state_dict
        self.device = device
        self._cache: Dict[tuple, torch.Tensor] = {}
        self._makers: Dict[tuple, object] = {}
        self._origin: Dict[int, tuple] = {}     # data_ptr of a kernel-ready tensor -> (kind, key)

    def has(self, key: str) -> bool:
        return key in self.sd

    def keys(self):
        return self.sd.keys()

    def _get(self, key: str) -> torch.Tensor:
        if key not in self.sd:
            raise KeyError(f"missing parameter {key}")
        return self.sd[key].detach().to(self.device)

    def _memo(self, kind, key, fn):
        k = (kind, key)
        if k not in self._cache:
            self._cache[k] = fn()
            self._makers[k] = fn
            self._origin[self._cache[k].data_ptr()] = k
        return self._cache[k]

    def origin(self, t: torch.Tensor) -> tuple:
        """(kind, key) of a kernel-ready tensor handed out by this store (training: where a weight gradient belongs)."""
        return self._origin[t.data_ptr()]

    def refresh(self) -> None:
        """Re-derive every kernel-ready tensor IN PLACE from the (updated) state dict: pointers and TMA descriptors built
        on them stay valid (training: after each optimizer step)."""
        for k, fn in self._makers.items():
            self._cache[k].copy_(fn())

    def f32(self, key: str) -> torch.Tensor:
        def make():
            t = self._get(key).to(F32).contiguous()
            # the kernels read these vectors 16 bytes at a time; a training state dict holds views into the optimizer's
            # flat master buffers at arbitrary element offsets
            return t.clone() if t.data_ptr() % 16 else t
        return self._memo("f32", key, make)

    def linear(self, key: str) -> torch.Tensor:
        """[N, K] bf16 — nn.Linear weights are already K-major."""
        return self._memo("lin", key, lambda: self._get(key).to(BF16).reshape(self.sd[key].shape[0], -1).contiguous())

    def conv3(self, key: str, cin_pad: Optional[int] = None) -> torch.Tensor:
        """[Cout, Cin, 3, 3] -> [Cout, 9*Cin_pad] bf16, K index = (ky*3 + kx)*Cin_pad + ci."""
        def make():
            w = self._get(key).to(F32)
            co, ci = w.shape[:2]
            cp = cin_pad or ci
            out = torch.zeros(co, 3, 3, cp, device=self.device, dtype=F32)
            out[..., :ci] = w.permute(0, 2, 3, 1)
            return out.reshape(co, 9 * cp).to(BF16).contiguous()
        return self._memo(("conv3", cin_pad), key, make)

    def conv3_direct(self, key: str) -> torch.Tensor:
        """[Cout, Cin, 3, 3] -> fp32 [3][3][Cin][Cout] for the direct small-channel conv."""
        return self._memo("conv3d", key, lambda: self._get(key).to(F32).permute(2, 3, 1, 0).contiguous())

    def tconv(self, key: str) -> torch.Tensor:
        """Conv3d (3,1,1) [C, C, 3, 1, 1] -> [C, 3*C] bf16, K index = kt*C + ci."""
        def make():
            w = self._get(key).to(F32)
            co, ci = w.shape[:2]
            return w.reshape(co, ci, 3).permute(0, 2, 1).reshape(co, 3 * ci).to(BF16).contiguous()
        return self._memo("tconv", key, make)

    def qkv(self, prefix: str) -> torch.Tensor:
        def make():
            return torch.cat([self._get(prefix + f"to_{n}.weight").to(BF16) for n in ("q", "k", "v")], 0).contiguous()
        return self._memo("qkv", prefix, make)

    def cat_rows(self, keys: List[str], kind: str) -> torch.Tensor:
        def make():
            dt = BF16 if kind == "bf16" else F32
            return torch.cat([self._get(k).to(dt).reshape(self.sd[k].shape[0], -1) if kind == "bf16"
                              else self._get(k).to(dt).reshape(-1) for k in keys], 0).contiguous()
        return self._memo(("cat", kind), tuple(keys), make)

    def cc_split(self, key: str, c_feat: int):
        """cc_projection.weight [C, C + 12] (controlnet_sdv_cam_infer.py:84) as its two column blocks, both bf16: the
        feature block [C, C] (a GEMM on the conditioning features) and the camera block [C, 12] (a per-frame row vector)."""
        feat = self._memo(("ccf", c_feat), key, lambda: self._get(key).to(F32)[:, :c_feat].to(BF16).contiguous())
        cam = self._memo(("ccc", c_feat), key, lambda: self._get(key).to(F32)[:, c_feat:].to(BF16).contiguous())
        return feat, cam

    def alpha(self, key: str) -> float:
        """sigmoid(mix_factor) of an AlphaBlender (image_only_indicator is all zeros on this path)."""
        return float(torch.sigmoid(self._get(key).to(F32).reshape(-1)[0]).item())


class Pool:
    """Reuses activation buffers between ops of one in-order op list."""

    def __init__(self, device, reuse: bool = True):
        self.device = device
        self.free = defaultdict(list)
        self.bytes = 0
        self.reuse = reuse   # False (training): every activation stays alive for the backward pass

    def get(self, rows: int, cols: int, dtype=BF16) -> torch.Tensor:
        key = (rows, cols, dtype)
        if self.free[key]:
            return self.free[key].pop()
        t = torch.empty(rows, cols, device=self.device, dtype=dtype)
        self.bytes += t.numel() * t.element_size()
        return t

    def put(self, *ts) -> None:
        if not self.reuse:
            return
        for t in ts:
            if t is not None and not getattr(t, "_pt_no_pool", False):   # peer-mapped exchange buffers are dedicated
                self.free[(t.shape[0], t.shape[1], t.dtype)].append(t)


class NetPlan:
    """Op lists of one network (kind 'unet' or 'controlnet') for a fixed problem shape."""

    def __init__(self, kind: str, cfg: SVDConfig, weights: WeightStore, *, batch: int, frames: int, height: int,
                 width: int, device, cam: bool = False, bbox: bool = False, cond_hw: Optional[tuple] = None,
                 sigmas: Optional[torch.Tensor] = None, step_index: Optional[torch.Tensor] = None,
                 x_in: Optional[torch.Tensor] = None, residual_bufs: Optional[List[torch.Tensor]] = None,
                 ctx_batch: Optional[int] = None, row_offset: int = 0, train: bool = False, defer_injection: bool = False):
        assert kind in ("unet", "controlnet")
        self.kind, self.cfg, self.w = kind, cfg, weights
        # train=True (posetraj_b200.train_engine, BASELINE configs[3]): no buffer reuse, pre-activations kept (GEGLU / SiLU as
        # separate passes, attention log-sum-exp written), parameter-dependent constants re-evaluated every step
        self.train = train
        # defer_injection (UNet only): the encoder does not read the ControlNet residuals — the skips are written without them
        # and `inject_ops` adds m_i * r_i afterwards — so that `step_ops[:split_index]` (time embedding, encoder, first half
        # of the mid block) can run CONCURRENTLY with the ControlNet on a second stream (pipeline.DenoiseEngine)
        self.defer_injection = defer_injection
        self.inject_ops: List = []
        self.split_index = 0
        self.alpha_updaters: List = []   # (mix_factor key, fn(alpha)) of every AlphaBlender baked into launch arguments
        self.B, self.F, self.H, self.W = batch, frames, height, width
        self.n = batch * frames
        # CFG-branch sharding: this plan computes rows [row_offset, row_offset + batch) of a call whose
        # full batch is ctx_batch.  Activations never cross rows, but the 1-token cross-attention constants do
        # (fact 11), so every shard keeps ALL ctx_batch image embeddings.
        self.ctx_B = ctx_batch if ctx_batch is not None else batch
        self.row_offset = row_offset
        if row_offset + batch > self.ctx_B:
            raise ValueError("row_offset + batch exceeds ctx_batch")
        self.device = device
        self.cam, self.bbox = cam, bbox
        self.pool = Pool(device, reuse=not train)
        self.step_ops: List = []    # every denoise step
        self.embed_ops: List = []   # when encoder_hidden_states / added_time_ids change
        self.cond_ops: List = []    # when controlnet_cond / camera_cond change (ControlNet only)
        self.scale_ops: List = []   # zero-conv GEMMs scaled by `conditioning_scale`
        # conditioning_scale as a DEVICE scalar: the zero-conv GEMMs read it when they run, so a captured CUDA graph
        # of the step follows the value of the current call (controlnet_sdv.py:641-643)
        self.cond_scale = torch.ones(1, device=device, dtype=F32)
        self._cond_scale_host = 1.0
        ch = cfg.block_out_channels
        if height % (2 ** (len(ch) - 1)) or width % (2 ** (len(ch) - 1)):
            raise ValueError("latent height/width must be divisible by 8 (three stride-2 levels)")

        # ---- inputs -------------------------------------------------------------------------------------
        P0 = (height + 1) * (width + 1)
        self.cin_pad = _pad64(cfg.in_channels)
        self.x_in = x_in if x_in is not None else torch.zeros(self.n * P0, self.cin_pad, device=device, dtype=BF16)
        self.ehs = torch.zeros(self.ctx_B, cfg.cross_attention_dim, device=device, dtype=F32)
        self.time_ids = torch.zeros(batch * 3, device=device, dtype=F32)
        self.t_buf = torch.zeros(batch, device=device, dtype=F32)
        self.sigmas, self.step_index = sigmas, step_index
        # GroupNorm partial-sum workspace: at most 4 CTAs per SM plus one per statistics group
        self.stats = torch.zeros((2 * self.n + 4 * ops.NUM_SMS + 64) * 64 + 1024, device=device, dtype=torch.float64)
        self.level_hw = [(height >> i, width >> i) for i in range(len(ch))]

        # ---- time embedding ops (first in every step) ---------------------------------------------------
        temb = cfg.temb_dim
        self.temb_keys = sorted(k for k in weights.keys() if k.endswith("time_emb_proj.weight"))
        self.temb_off, off = {}, 0
        for k in self.temb_keys:
            self.temb_off[k[: -len(".weight")]] = off
            off += weights.sd[k].shape[0]
        self.tproj = torch.zeros(batch, off, device=device, dtype=F32)
        self._time_embedding_ops(temb)

        # ---- residual hand-off buffers ------------------------------------------------------------------
        self.res_shapes = self._residual_shapes()
        if residual_bufs is not None:
            self.res = residual_bufs
        else:
            self.res = [torch.zeros(r, c, device=device, dtype=BF16) for (r, c) in self.res_shapes]

        # ---- body ---------------------------------------------------------------------------------------
        if kind == "controlnet":
            self._build_controlnet(cond_hw)
        else:
            self._build_unet()

    # ================================================================================================
    # helpers
    # ================================================================================================
    def _residual_shapes(self):
        cfg = self.cfg
        ch = cfg.block_out_channels
        shapes = [(self.n * self.H * self.W, ch[0])]
        n = len(ch)
        for i in range(n):
            h, w = self.level_hw[i]
            for _ in range(cfg.layers_per_block):
                shapes.append((self.n * h * w, ch[i]))
            if i < n - 1:
                h2, w2 = self.level_hw[i + 1]
                shapes.append((self.n * h2 * w2, ch[i]))
        h, w = self.level_hw[-1]
        shapes.append((self.n * h * w, ch[-1]))  # mid
        return shapes

    def _time_embedding_ops(self, temb: int) -> None:
        cfg, w, dev, B = self.cfg, self.w, self.device, self.B
        c0 = cfg.block_out_channels[0]
        self.t_sin = torch.zeros(B, c0, device=dev, dtype=F32)
        self.t_h = torch.zeros(B, temb, device=dev, dtype=F32)
        self.emb = torch.zeros(B, temb, device=dev, dtype=F32)
        self.a_sin = torch.zeros(B * 3, cfg.addition_time_embed_dim, device=dev, dtype=F32)
        self.a_h = torch.zeros(B, temb, device=dev, dtype=F32)
        if self.sigmas is not None:
            self.step_ops.append(ops.SinCos(self.t_sin, sigmas=self.sigmas, step_index=self.step_index, name="time_proj"))
        else:
            self.step_ops.append(ops.SinCos(self.t_sin, t=self.t_buf, name="time_proj"))
        # training keeps the pre-activation: SiLU moves from the producer's output to the consumer's input (same function)
        tr = self.train
        self.step_ops.append(ops.SmallLinear(self.t_sin, w.linear("time_embedding.linear_1.weight"), self.t_h,
                                             w.f32("time_embedding.linear_1.bias"), act_out_silu=not tr, name="time_embedding.1"))
        self.step_ops.append(ops.SmallLinear(self.t_h, w.linear("time_embedding.linear_2.weight"), self.emb,
                                             w.f32("time_embedding.linear_2.bias"), act_in_silu=tr, name="time_embedding.2"))
        # added_time_ids: Timesteps(256) of the flattened ids -> [B, 768] -> add_embedding (controlnet_sdv.py:577-581)
        self.step_ops.append(ops.SinCos(self.a_sin, t=self.time_ids, name="add_time_proj"))
        a_in = self.a_sin.view(B, 3 * cfg.addition_time_embed_dim)
        self.step_ops.append(ops.SmallLinear(a_in, w.linear("add_embedding.linear_1.weight"), self.a_h,
                                             w.f32("add_embedding.linear_1.bias"), act_out_silu=not tr, name="add_embedding.1"))
        self.step_ops.append(ops.SmallLinear(self.a_h, w.linear("add_embedding.linear_2.weight"), self.emb,
                                             w.f32("add_embedding.linear_2.bias"), act_in_silu=tr, accumulate=True, name="add_embedding.2"))
        # all time_emb_proj(silu(emb)) of the network in one GEMV
        w_all = w.cat_rows(self.temb_keys, "bf16")
        b_all = w.cat_rows([k[: -len("weight")] + "bias" for k in self.temb_keys], "f32")
        self.step_ops.append(ops.SmallLinear(self.emb, w_all, self.tproj, b_all, act_in_silu=True, name="time_emb_proj*"))

    def _tvec(self, prefix: str, cout: int) -> torch.Tensor:
        off = self.temb_off[prefix + "time_emb_proj"]
        return self.tproj[:, off: off + cout]

    def _gn(self, x0, x1, key, *, rows_per_stat, eps, silu, halo=None) -> torch.Tensor:
        Cc = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
        if halo is not None:
            n_img = x0.shape[0] // (halo[0] * halo[1])
            out = self.pool.get(n_img * (halo[0] + 1) * (halo[1] + 1), Cc)
        else:
            out = self.pool.get(x0.shape[0], Cc)
        self.step_ops.append(ops.GroupNorm(x0, out, self.w.f32(key + ".weight"), self.w.f32(key + ".bias"), self.stats,
                                           rows_per_stat=rows_per_stat, eps=eps, silu=silu, x1=x1, halo=halo, name=key))
        return out

    def _ln(self, x, key, **kw) -> torch.Tensor:
        out = self.pool.get(x.shape[0], x.shape[1])
        self.step_ops.append(ops.LayerNorm(x, out, self.w.f32(key + ".weight"), self.w.f32(key + ".bias"), name=key, **kw))
        return out

    def _gemm(self, a0, wt, n_out_cols, *, out=None, name="gemm", **kw) -> torch.Tensor:
        if out is None:
            rows = kw.pop("out_rows", a0.shape[0])
            out = self.pool.get(rows, n_out_cols)
        else:
            kw.pop("out_rows", None)
        self.step_ops.append(ops.Gemm(a0, wt, out, name=name, **kw))
        return out

    def _ff(self, x, prefix: str, *, res1, res1_scale: float = 1.0, res2=None, res2_scale: float = 1.0,
            acc_scale: float = 1.0, name: str) -> torch.Tensor:
        """FeedForward (GEGLU proj -> Linear) + residual terms.  Widths whose output accumulator fits TMEM next to the
        hidden one (C <= 320: level 0) run as ONE kernel (ops.FusedMlp) and never materialise the [rows, 4C] hidden
        tensor; wider levels keep the GEGLU GEMM + output GEMM pair."""
        w = self.w
        Cc = x.shape[1]
        w1, b1 = w.linear(prefix + "net.0.proj.weight"), w.f32(prefix + "net.0.proj.bias")
        w2, b2 = w.linear(prefix + "net.2.weight"), w.f32(prefix + "net.2.bias")
        if self.train:
            f0 = self._gemm(x, w1, 8 * Cc, bias=b1, name=name + ".proj")
            f1 = self.pool.get(x.shape[0], 4 * Cc)
            self.step_ops.append(ops.GegluFwd(f0, f1, name=name + ".geglu"))
            return self._gemm(f1, w2, Cc, bias=b2, acc_scale=acc_scale, res1=res1, res1_scale=res1_scale, res2=res2,
                              res2_scale=res2_scale, name=name + ".out")
        if FUSED_MLP and ops.FusedMlp.supported(Cc) and w2.shape == (Cc, 4 * Cc):
            out = self.pool.get(x.shape[0], Cc)
            self.step_ops.append(ops.FusedMlp(x, w1, b1, w2, b2, out, acc_scale=acc_scale, res1=res1, res1_scale=res1_scale,
                                              res2=res2, res2_scale=res2_scale, name=name + ".mlp"))
            return out
        f1 = self._gemm(x, w1, 4 * Cc, geglu=True, bias=b1, name=name + ".geglu")
        out = self._gemm(f1, w2, Cc, bias=b2, acc_scale=acc_scale, res1=res1, res1_scale=res1_scale, res2=res2,
                         res2_scale=res2_scale, name=name + ".out")
        self.pool.put(f1)
        return out

    # ================================================================================================
    # blocks
    # ================================================================================================
    def resblock(self, prefix: str, x0, x1, cout: int, hw: tuple, eps: float, *, out2=None, aux=None,
                 aux_scale: float = 0.0, res2=None) -> torch.Tensor:
        """SpatioTemporalResBlock. x1 is the skip half of an un-materialised channel concat."""
        w, B, Fr = self.w, self.B, self.F
        H, W = hw
        HW = H * W
        rows = self.n * HW
        cin = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
        s, t = prefix + "spatial_res_block.", prefix + "temporal_res_block."
        taps = ops.conv3x3_taps(W)
        # spatial ResnetBlock2D
        g1 = self._gn(x0, x1, s + "norm1", rows_per_stat=HW, eps=eps, silu=True, halo=hw)
        h1 = self._gemm(g1, w.conv3(s + "conv1.weight"), cout, taps=taps, bias=w.f32(s + "conv1.bias"),
                        rowvec=self._tvec(s, cout), rowvec_mode=1, rv=(Fr * HW, 1, 1), halo=hw, out_rows=rows,
                        name=s + "conv1")
        self.pool.put(g1)
        g2 = self._gn(h1, None, s + "norm2", rows_per_stat=HW, eps=eps, silu=True, halo=hw)
        self.pool.put(h1)
        if cin != cout:
            sc = self._gemm(x0, w.linear(s + "conv_shortcut.weight"), cout, a1=x1, bias=w.f32(s + "conv_shortcut.bias"),
                            name=s + "conv_shortcut")
        else:
            assert x1 is None
            sc = x0
        xs = self._gemm(g2, w.conv3(s + "conv2.weight"), cout, taps=taps, bias=w.f32(s + "conv2.bias"), res1=sc,
                        halo=hw, out_rows=rows, name=s + "conv2")
        self.pool.put(g2)
        if sc is not x0:
            self.pool.put(sc)
        # TemporalResnetBlock on the same rows viewed [B, F*HW, C]; 5-D GroupNorm statistics per batch row
        t1 = self._gn(xs, None, t + "norm1", rows_per_stat=Fr * HW, eps=eps, silu=True)
        t2 = self._gemm(t1, w.tconv(t + "conv1.weight"), cout, batches=B, taps=(-HW, 0, HW), bias=w.f32(t + "conv1.bias"),
                        rowvec=self._tvec(t, cout), rowvec_mode=1, rv=(Fr * HW, 1, 1), name=t + "conv1")
        self.pool.put(t1)
        t3 = self._gn(t2, None, t + "norm2", rows_per_stat=Fr * HW, eps=eps, silu=True)
        self.pool.put(t2)
        alpha = w.alpha(prefix + "time_mixer.mix_factor")
        # blend(xs, xs + conv2(.)) = xs + (1 - alpha) * conv2(.)
        out = self._gemm(t3, w.tconv(t + "conv2.weight"), cout, batches=B, taps=(-HW, 0, HW), bias=w.f32(t + "conv2.bias"),
                         acc_scale=1.0 - alpha, res1=xs, res2=res2, res2_scale=1.0, out2=out2, aux=aux,
                         aux_scale=aux_scale, name=t + "conv2")
        blend = self.step_ops[-1]
        blend.io.mix = (prefix + "time_mixer.mix_factor", xs, alpha)   # out = alpha*xs + (1-alpha)*(xs + conv2): see train_engine

        def upd(a, op=blend):
            op.args.acc_scale = op.io.acc_scale = 1.0 - a
            op.io.mix = (op.io.mix[0], op.io.mix[1], a)
        self.alpha_updaters.append((prefix + "time_mixer.mix_factor", upd))
        self.pool.put(t3, xs)
        return out

    def _xvec(self, attn_prefix: str, Cc: int) -> torch.Tensor:
        """Constant of the degenerate 1-token cross-attention: to_out(to_v(e_b)) + bias, fp32 [B, C]."""
        w, dev, B = self.w, self.device, self.ctx_B
        tmp = torch.zeros(B, Cc, device=dev, dtype=F32)
        vec = torch.zeros(B, Cc, device=dev, dtype=F32)
        self.embed_ops.append(ops.SmallLinear(self.ehs, w.linear(attn_prefix + "to_v.weight"), tmp, None,
                                              name=attn_prefix + "to_v"))
        self.embed_ops.append(ops.SmallLinear(tmp, w.linear(attn_prefix + "to_out.0.weight"), vec,
                                              w.f32(attn_prefix + "to_out.0.bias"), name=attn_prefix + "to_out"))
        return vec

    def _frame_pos_emb(self, prefix: str, Cc: int) -> torch.Tensor:
        """time_pos_embed(time_proj(arange(F))) — step-invariant, evaluated once at build (modified_svd.py:168-179)."""
        w, dev, Fr = self.w, self.device, self.F
        sp = torch.cuda.current_stream().cuda_stream
        sin = torch.zeros(Fr, Cc, device=dev, dtype=F32)
        hid = torch.zeros(Fr, 4 * Cc, device=dev, dtype=F32)
        out = torch.zeros(Fr, Cc, device=dev, dtype=F32)
        frames = torch.arange(Fr, device=dev, dtype=F32)
        tr = self.train
        lst = [ops.SinCos(sin, t=frames, name=prefix + "time_proj"),
               ops.SmallLinear(sin, w.linear(prefix + "time_pos_embed.linear_1.weight"), hid,
                               w.f32(prefix + "time_pos_embed.linear_1.bias"), act_out_silu=not tr, name=prefix + "time_pos_embed.1"),
               ops.SmallLinear(hid, w.linear(prefix + "time_pos_embed.linear_2.weight"), out,
                               w.f32(prefix + "time_pos_embed.linear_2.bias"), act_in_silu=tr, name=prefix + "time_pos_embed.2")]
        for op in lst:
            op.launch(sp)
        if tr:   # the MLP is trained: re-evaluated with the other parameter-dependent constants every step
            self.embed_ops += lst
        return out

    def transformer(self, prefix: str, x, heads: int, hw: tuple, *, out2=None, aux=None, aux_scale: float = 0.0):
        """TransformerSpatioTemporalModel."""
        w, B, Fr = self.w, self.B, self.F
        H, W = hw
        HW = H * W
        Cc = x.shape[1]
        sb, tb = prefix + "transformer_blocks.0.", prefix + "temporal_transformer_blocks.0."
        xvec_s = self._xvec(sb + "attn2.", Cc)[self.row_offset: self.row_offset + B]
        xvec_t = self._xvec(tb + "attn2.", Cc)
        rv_t = (Fr * HW, HW, self.ctx_B)
        if self.ctx_B != B:
            # hidden row (b, s) of the FULL batch gets the vector of batch ((b*HW + s) mod ctx_B); with only rows
            # [row_offset, ...) present the kernel sees local b, so pre-rotate the table by row_offset*HW instead
            if B != 1:
                raise ValueError("row sharding of the temporal context is implemented for one row per shard")
            rot = torch.zeros_like(xvec_t)
            src, shift, nb = xvec_t, (self.row_offset * HW) % self.ctx_B, self.ctx_B
            self.embed_ops.append(ops.TorchOp(lambda s=src, d=rot, k=shift: d.copy_(torch.roll(s, -k, 0)),
                                              name=tb + "attn2.rotate"))
            xvec_t = rot
        pos = self._frame_pos_emb(prefix, Cc)
        g = self._gn(x, None, prefix + "norm", rows_per_stat=HW, eps=1e-6, silu=False)
        h = self._gemm(g, w.linear(prefix + "proj_in.weight"), Cc, bias=w.f32(prefix + "proj_in.bias"), name=prefix + "proj_in")
        self.pool.put(g)
        # --- spatial BasicTransformerBlock
        l1 = self._ln(h, sb + "norm1")
        qkv = self._gemm(l1, w.qkv(sb + "attn1."), 3 * Cc, name=sb + "attn1.qkv")
        self.pool.put(l1)
        att = self.pool.get(x.shape[0], Cc)
        lse = torch.empty(self.n * heads * HW, device=self.device, dtype=F32) if self.train else None
        self.step_ops.append(ops.AttnSpatial(qkv, att, n_img=self.n, heads=heads, name=sb + "attn1", lse=lse))
        self.pool.put(qkv)
        # h2 = attn1 + h, then + attn2 (constant per batch row): both in one epilogue
        h2 = self._gemm(att, w.linear(sb + "attn1.to_out.0.weight"), Cc, bias=w.f32(sb + "attn1.to_out.0.bias"), res1=h,
                        rowvec=xvec_s, rowvec_mode=1, rv=(Fr * HW, 1, 1), name=sb + "attn1.to_out")
        self.pool.put(att, h)
        l3 = self._ln(h2, sb + "norm3")
        h3 = self._ff(l3, sb + "ff.", res1=h2, name=sb + "ff")
        self.pool.put(l3, h2)
        # --- TemporalBasicTransformerBlock on (h3 + frame position embedding)
        ht = self.pool.get(x.shape[0], Cc)
        l_in = self._ln(h3, tb + "norm_in", addvec=pos, hw=HW, frames=Fr, sum_out=ht)
        t1 = self._ff(l_in, tb + "ff_in.", res1=ht, name=tb + "ff_in")
        self.pool.put(l_in, ht)
        l1t = self._ln(t1, tb + "norm1")
        qkv_t = self._gemm(l1t, w.qkv(tb + "attn1."), 3 * Cc, name=tb + "attn1.qkv")
        self.pool.put(l1t)
        att_t = self.pool.get(x.shape[0], Cc)
        self.step_ops.append(ops.AttnTemporal(qkv_t, att_t, batch=B, frames=Fr, hw=HW, heads=heads, name=tb + "attn1"))
        self.pool.put(qkv_t)
        t2 = self._gemm(att_t, w.linear(tb + "attn1.to_out.0.weight"), Cc, bias=w.f32(tb + "attn1.to_out.0.bias"), res1=t1,
                        rowvec=xvec_t, rowvec_mode=2, rv=rv_t, name=tb + "attn1.to_out")
        self.pool.put(att_t, t1)
        l3t = self._ln(t2, tb + "norm3")
        alpha = w.alpha(prefix + "time_mixer.mix_factor")
        # blend: alpha*h3 + (1-alpha)*(ff(.) + t2)
        hb = self._ff(l3t, tb + "ff.", acc_scale=1.0 - alpha, res1=t2, res1_scale=1.0 - alpha, res2=h3, res2_scale=alpha,
                      name=tb + "ff+mix")
        blend = self.step_ops[-1]
        if hasattr(blend, "io") and hasattr(blend.io, "mix"):   # (the fused-MLP kernel of the inference plans has no record)
            blend.io.mix = (prefix + "time_mixer.mix_factor", h3, alpha)   # out = alpha*h3 + (1-alpha)*(ff + t2)

        def upd(a, op=blend):
            op.args.acc_scale, op.args.res1_scale, op.args.res2_scale = 1.0 - a, 1.0 - a, a
            if hasattr(op, "io") and hasattr(op.io, "mix"):
                op.io.acc_scale, op.io.res1_scale, op.io.res2_scale = 1.0 - a, 1.0 - a, a
                op.io.mix = (op.io.mix[0], op.io.mix[1], a)
        self.alpha_updaters.append((prefix + "time_mixer.mix_factor", upd))
        self.pool.put(l3t, t2, h3)
        out = self._gemm(hb, w.linear(prefix + "proj_out.weight"), Cc, bias=w.f32(prefix + "proj_out.bias"), res1=x,
                         out2=out2, aux=aux, aux_scale=aux_scale, name=prefix + "proj_out")
        self.pool.put(hb)
        return out

    def downsample(self, key: str, x, hw: tuple, **kw) -> torch.Tensor:
        H, W = hw
        Cc = x.shape[1]
        xh = self.pool.get(self.n * (H + 1) * (W + 1), Cc)
        self.step_ops.append(ops.Upsample2x(x, xh, n=self.n, H=H, W=W, halo=True, scale=1, name=key + ".halo"))
        out = self._gemm(xh, self.w.conv3(key + ".weight"), Cc, taps=ops.conv3x3_taps(W), bias=self.w.f32(key + ".bias"),
                         halo=hw, ostride=2, out_rows=self.n * (H // 2) * (W // 2), name=key, **kw)
        self.pool.put(xh)
        return out

    def upsample(self, key: str, x, hw: tuple) -> torch.Tensor:
        H, W = hw
        Cc = x.shape[1]
        xh = self.pool.get(self.n * (2 * H + 1) * (2 * W + 1), Cc)
        self.step_ops.append(ops.Upsample2x(x, xh, n=self.n, H=H, W=W, halo=True, scale=2, name=key + ".nearest2x"))
        out = self._gemm(xh, self.w.conv3(key + ".weight"), Cc, taps=ops.conv3x3_taps(2 * W), bias=self.w.f32(key + ".bias"),
                         halo=(2 * H, 2 * W), out_rows=self.n * 4 * H * W, name=key)
        self.pool.put(xh)
        return out

    # ================================================================================================
    # trunks
    # ================================================================================================
    def _encoder(self, on_skip, conv_in_res=None):
        """conv_in + down blocks + mid block.  `on_skip(i, kwargs)` decorates the GEMM producing skip i and
        `on_skip.done(i, tensor)` is told about the result; returns the mid-block output."""
        cfg, w = self.cfg, self.w
        ch, heads = cfg.block_out_channels, cfg.num_attention_heads
        n = len(ch)
        hw = self.level_hw[0]
        kw = on_skip(0)
        x = self._gemm(self.x_in, w.conv3("conv_in.weight", self.cin_pad), ch[0], taps=ops.conv3x3_taps(hw[1]),
                       bias=w.f32("conv_in.bias"), res1=conv_in_res, halo=hw, out_rows=self.n * hw[0] * hw[1],
                       name="conv_in", alg_k=9 * cfg.in_channels, **kw)
        on_skip.done(0, x)
        idx = 1
        for i in range(n):
            hw = self.level_hw[i]
            for j in range(cfg.layers_per_block):
                p = f"down_blocks.{i}."
                eps = 1e-6 if i < n - 1 else 1e-5
                if i < n - 1:
                    y = self.resblock(p + f"resnets.{j}.", x, None, ch[i], hw, eps)
                    on_skip.release(x)
                    kw = on_skip(idx)
                    x = self.transformer(p + f"attentions.{j}.", y, heads[i], hw, **kw)
                    self.pool.put(y)
                else:
                    kw = on_skip(idx)
                    y = self.resblock(p + f"resnets.{j}.", x, None, ch[i], hw, eps, **kw)
                    on_skip.release(x)
                    x = y
                on_skip.done(idx, x)
                idx += 1
            if i < n - 1:
                kw = on_skip(idx)
                y = self.downsample(f"down_blocks.{i}.downsamplers.0.conv", x, hw, **kw)
                on_skip.release(x)
                x = y
                on_skip.done(idx, x)
                idx += 1
        # mid block: res(1e-5) -> transformer -> res(1e-5)
        hw = self.level_hw[-1]
        y = self.resblock("mid_block.resnets.0.", x, None, ch[-1], hw, 1e-5)
        on_skip.release(x)
        z = self.transformer("mid_block.attentions.0.", y, heads[-1], hw)
        self.pool.put(y)
        return z

    def _build_controlnet(self, cond_hw):
        cfg, w, dev = self.cfg, self.w, self.device
        ch = cfg.block_out_channels
        plan = self

        # ---- conditioning embedding (step-invariant: its own op list) ----------------------------------
        Hc, Wc = cond_hw if cond_hw is not None else (self.H * 8, self.W * 8)
        if (Hc // 8, Wc // 8) != (self.H, self.W):
            raise ValueError("controlnet_cond must be 8x the latent resolution")
        self.cond_hw = (Hc, Wc)
        self.cond_in = torch.zeros(self.n, cfg.conditioning_channels, Hc, Wc, device=dev, dtype=F32)
        self.cond_in2 = torch.zeros_like(self.cond_in) if self.bbox else None
        self.cam_in = torch.zeros(self.n, 12, device=dev, dtype=F32) if self.cam else None
        self.use_cam = False
        self.cond_emb = torch.zeros(self.n * self.H * self.W, ch[0], device=dev, dtype=BF16)
        self._build_cond_embedding()

        class Skips:
            """ControlNet: every skip goes through its zero-conv right away (controlnet_sdv.py:632-638)."""

            def __call__(self_, i):
                return {}

            def done(self_, i, t):
                key = f"controlnet_down_blocks.{i}"
                g = ops.Gemm(t, w.linear(key + ".weight"), plan.res[i], bias=w.f32(key + ".bias"), acc_scale=1.0,
                             acc_scale_dev=plan.cond_scale, name=key)
                plan.step_ops.append(g)
                plan.scale_ops.append(g)

            def release(self_, t):
                plan.pool.put(t)

        z = self._encoder(Skips(), conv_in_res=self.cond_emb)
        hw = self.level_hw[-1]
        mid = self.resblock("mid_block.resnets.1.", z, None, ch[-1], hw, 1e-5)
        self.pool.put(z)
        g = ops.Gemm(mid, w.linear("controlnet_mid_block.weight"), self.res[-1], bias=w.f32("controlnet_mid_block.bias"),
                     acc_scale=1.0, acc_scale_dev=self.cond_scale, name="controlnet_mid_block")
        self.step_ops.append(g)
        self.scale_ops.append(g)
        self.pool.put(mid)

    def _build_cond_embedding(self):
        """ControlNetConditioningEmbeddingSVD[_CAM] (controlnet_sdv.py:95-116, controlnet_sdv_cam_infer.py:96-122,
        controlnet_sdv_bbox.py:109-138): 3->16->16->32(s2)->32->96(s2)->96->256(s2) with SiLU, then conv_out."""
        cfg, w, dev = self.cfg, self.w, self.device
        ce = cfg.conditioning_embedding_out_channels
        Hc, Wc = self.cond_hw
        p = "controlnet_cond_embedding."
        self.cam_rowvec = None
        if self.train:
            return self._build_cond_embedding_train()
        towers = [("conv_in", "blocks", self.cond_in)]
        if self.bbox:
            towers.append(("conv_in_2", "blocks_2", self.cond_in2))
        feats = []
        for cin_name, blocks_name, src in towers:
            layers = [(cin_name, cfg.conditioning_channels, ce[0], 1)]
            for i in range(len(ce) - 1):
                layers.append((f"{blocks_name}.{2 * i}", ce[i], ce[i], 1))
                layers.append((f"{blocks_name}.{2 * i + 1}", ce[i], ce[i + 1], 2))
            x, xh_w = src, (Hc, Wc)
            x_is_nchw, x_halo = True, False
            for li, (name, cin, cout, stride) in enumerate(layers):
                H, W = xh_w
                oH, oW = H // stride, W // stride
                last = li == len(layers) - 1
                nxt_direct = (not last) and layers[li + 1][1] <= 16 and layers[li + 1][2] <= 32
                if cin <= 16 and cout <= 32:
                    # narrow layers: direct conv; output compact for a direct consumer, zero-haloed for a GEMM consumer
                    out_halo = not nxt_direct
                    rows = self.n * ((oH + 1) * (oW + 1) if out_halo else oH * oW)
                    out = torch.zeros(rows, _pad64(cout) if out_halo else cout, device=dev, dtype=BF16)
                    self.cond_ops.append(ops.ConvDirect(x, w.conv3_direct(p + name + ".weight"), w.f32(p + name + ".bias"), out,
                                                        n=self.n, H=H, W=W, cin=cin, cout=cout, stride=stride, silu=True,
                                                        in_nchw_f32=x_is_nchw, out_halo=out_halo, name=p + name))
                else:
                    assert x_halo
                    out = torch.zeros(self.n * (oH + 1) * (oW + 1), _pad64(cout), device=dev, dtype=BF16)
                    self.cond_ops.append(ops.Gemm(x, w.conv3(p + name + ".weight", _pad64(cin)), out, taps=ops.conv3x3_taps(W),
                                                  n_out=cout, bias=w.f32(p + name + ".bias"), halo=(H, W), ostride=stride,
                                                  out_halo=True, act_silu=True, name=p + name))
                    out_halo = True
                x, xh_w, x_is_nchw, x_halo = out, (oH, oW), False, out_halo
            feats.append(x)
        c_last = ce[-1]
        for ti, feat in enumerate(feats):
            if ti == 0 and self.cam:
                # cc_projection on [features | camera]: features through the GEMM, the 12 camera columns as a
                # per-frame row vector (controlnet_sdv_cam_infer.py:109-119)
                wfull = w._get(p + "cc_projection.weight").to(F32)
                self._cc_w_feat = wfull[:, :c_last].to(BF16).contiguous()
                self._cc_w_cam = wfull[:, c_last:].to(BF16).contiguous()
                self.cam_rowvec = torch.zeros(self.n, c_last, device=dev, dtype=F32)
                self.cam_ops = [ops.SmallLinear(self.cam_in, self._cc_w_cam, self.cam_rowvec, w.f32(p + "cc_projection.bias"),
                                                name=p + "cc_projection.cam")]
                proj = torch.zeros_like(feat)
                H, W = self.H, self.W
                self.cam_gemm = ops.Gemm(feat, self._cc_w_feat, proj, taps=(0,), n_out=c_last, rowvec=self.cam_rowvec,
                                         rowvec_mode=1, rv=((H + 1) * (W + 1), 1, 1), halo=(H, W), out_halo=True,
                                         name=p + "cc_projection")
                self.feat_plain, self.feat_cam = feat, proj
            # conv_out (zero-init in a fresh ControlNet); the bbox tower is projected with the SAME conv_out (bbox.py:134)
            H, W = self.H, self.W
            self._cond_out_ops = getattr(self, "_cond_out_ops", [])
            kw = dict(taps=ops.conv3x3_taps(W), bias=w.f32(p + "conv_out.bias"), halo=(H, W), name=p + "conv_out")
            if ti == 0:
                self.cond_out_plain = ops.Gemm(feat, w.conv3(p + "conv_out.weight"), self.cond_emb, **kw)
                if self.cam:
                    self.cond_out_cam = ops.Gemm(self.feat_cam, w.conv3(p + "conv_out.weight"), self.cond_emb, **kw)
            else:
                self.cond_out_bbox = ops.Gemm(feat, w.conv3(p + "conv_out.weight"), self.cond_emb, res1=self.cond_emb, **kw)

    def _build_cond_embedding_train(self):
        """Training variant of the conditioning embedding: EVERY conv through the implicit-GEMM kernel on the zero-haloed
        token layout (channels padded to 64), SiLU as its own pass so that the pre-activations exist for the backward."""
        cfg, w, dev = self.cfg, self.w, self.device
        ce = cfg.conditioning_embedding_out_channels
        Hc, Wc = self.cond_hw
        p = "controlnet_cond_embedding."
        towers = [("conv_in", "blocks", self.cond_in)]
        if self.bbox:
            towers.append(("conv_in_2", "blocks_2", self.cond_in2))
        feats = []
        for cin_name, blocks_name, src in towers:
            layers = [(cin_name, cfg.conditioning_channels, ce[0], 1)]
            for i in range(len(ce) - 1):
                layers.append((f"{blocks_name}.{2 * i}", ce[i], ce[i], 1))
                layers.append((f"{blocks_name}.{2 * i + 1}", ce[i], ce[i + 1], 2))
            x = torch.zeros(self.n * (Hc + 1) * (Wc + 1), _pad64(cfg.conditioning_channels), device=dev, dtype=BF16)
            self.cond_ops.append(ops.Layout(src, x, to_tokens=True, halo=True, name=p + cin_name + ".layout"))
            H, W = Hc, Wc
            for name, cin, cout, stride in layers:
                oH, oW = H // stride, W // stride
                z = torch.zeros(self.n * (oH + 1) * (oW + 1), _pad64(cout), device=dev, dtype=BF16)
                self.cond_ops.append(ops.Gemm(x, w.conv3(p + name + ".weight", _pad64(cin)), z, taps=ops.conv3x3_taps(W), n_out=cout,
                                              bias=w.f32(p + name + ".bias"), halo=(H, W), ostride=stride, out_halo=True,
                                              name=p + name))
                a = torch.zeros_like(z)
                self.cond_ops.append(ops.SiluFwd(z, a, name=p + name + ".silu"))
                x, H, W = a, oH, oW
            feats.append(x)
        H, W = self.H, self.W
        kw = dict(taps=ops.conv3x3_taps(W), bias=w.f32(p + "conv_out.bias"), halo=(H, W), name=p + "conv_out")
        self.cond_out_plain = ops.Gemm(feats[0], w.conv3(p + "conv_out.weight"), self.cond_emb, **kw)
        if self.cam:
            # cc_projection on [features | camera] (controlnet_sdv_cam_infer.py:109-119): the feature block as a GEMM, the
            # 12 camera columns as a per-frame row vector added in its epilogue
            c_last = ce[-1]
            w_feat, w_cam = w.cc_split(p + "cc_projection.weight", c_last)
            self.cam_rowvec = torch.zeros(self.n, c_last, device=dev, dtype=F32)
            self.cam_ops = [ops.SmallLinear(self.cam_in, w_cam, self.cam_rowvec, w.f32(p + "cc_projection.bias"),
                                            name=p + "cc_projection.cam")]
            proj = torch.zeros_like(feats[0])
            self.cam_gemm = ops.Gemm(feats[0], w_feat, proj, taps=(0,), n_out=c_last, rowvec=self.cam_rowvec, rowvec_mode=1,
                                     rv=((H + 1) * (W + 1), 1, 1), halo=(H, W), out_halo=True, name=p + "cc_projection")
            self.cond_out_cam = ops.Gemm(proj, w.conv3(p + "conv_out.weight"), self.cond_emb, **kw)
        if self.bbox:
            self.cond_out_bbox = ops.Gemm(feats[1], w.conv3(p + "conv_out.weight"), self.cond_emb, res1=self.cond_emb, **kw)

    def cond_op_list(self, use_cam: bool, use_bbox: bool) -> List:
        lst = list(self.cond_ops)
        if use_cam:
            if not self.cam:
                raise ValueError("camera_cond given but the model has no cc_projection")
            lst += self.cam_ops + [self.cam_gemm, self.cond_out_cam]
        else:
            lst.append(self.cond_out_plain)
        if use_bbox and self.bbox:
            lst.append(self.cond_out_bbox)
        return lst

    def _build_unet(self):
        cfg, w, dev = self.cfg, self.w, self.device
        ch, heads = cfg.block_out_channels, cfg.num_attention_heads
        mult = residual_multipliers(cfg)
        plan = self
        skips: Dict[int, torch.Tensor] = {}

        class Skips:
            """UNet: skip_i = h_i + m_i * r_i, written as a second output of the GEMM producing h_i."""

            def __call__(self_, i):
                rows, cols = plan.res_shapes[i]
                skips[i] = torch.empty(rows, cols, device=dev, dtype=BF16)
                if plan.defer_injection:
                    plan.inject_ops.append(ops.Axpy(skips[i], plan.res[i], skips[i], float(mult[i]), name=f"inject.{i}"))
                    return dict(out2=skips[i], aux=None, aux_scale=0.0)
                return dict(out2=skips[i], aux=plan.res[i], aux_scale=float(mult[i]))

            def done(self_, i, t):
                pass

            def release(self_, t):
                plan.pool.put(t)

        z = self._encoder(Skips())
        self.split_index = len(self.step_ops)
        hw = self.level_hw[-1]
        # second mid resnet + mid residual (unet...:469) as the second residual operand of its last GEMM
        x = self.resblock("mid_block.resnets.1.", z, None, ch[-1], hw, 1e-5, res2=self.res[-1])
        self.pool.put(z)
        n = len(ch)
        skip_idx = len(mult) - 1
        for i, (out_c, layers, has_attn, add_up) in enumerate(up_block_plan(cfg)):
            lvl = n - 1 - i
            hw = self.level_hw[lvl]
            for j, (res_in, skip_c) in enumerate(layers):
                sk = skips.pop(skip_idx)
                skip_idx -= 1
                assert x.shape[1] == res_in and sk.shape[1] == skip_c, (x.shape, sk.shape, res_in, skip_c)
                y = self.resblock(f"up_blocks.{i}.resnets.{j}.", x, sk, out_c, hw, 1e-6)
                self.pool.put(x)
                if has_attn:
                    x = self.transformer(f"up_blocks.{i}.attentions.{j}.", y, heads[lvl], hw)
                    self.pool.put(y)
                else:
                    x = y
            if add_up:
                y = self.upsample(f"up_blocks.{i}.upsamplers.0.conv", x, hw)
                self.pool.put(x)
                x = y
        hw = self.level_hw[0]
        g = self._gn(x, None, "conv_norm_out", rows_per_stat=hw[0] * hw[1], eps=1e-5, silu=True, halo=hw)
        self.pool.put(x)
        self.noise_pred = torch.zeros(self.n * hw[0] * hw[1], cfg.out_channels, device=dev, dtype=BF16)
        self.step_ops.append(ops.Gemm(g, w.conv3("conv_out.weight"), self.noise_pred, taps=ops.conv3x3_taps(hw[1]),
                                      bias=w.f32("conv_out.bias"), halo=hw, block_n=32, name="conv_out"))
        self.pool.put(g)

    # ================================================================================================
    # execution
    # ================================================================================================
    def set_conditioning_scale(self, scale: float) -> None:
        """Stream-ordered write of the device scalar every zero-conv GEMM multiplies its accumulator by."""
        if float(scale) != self._cond_scale_host:
            self.cond_scale.fill_(float(scale))
            self._cond_scale_host = float(scale)

    @staticmethod
    def run(op_list, stream_ptr: Optional[int] = None) -> None:
        if stream_ptr is None:
            stream_ptr = torch.cuda.current_stream().cuda_stream
        for op in op_list:
            op.launch(stream_ptr)
