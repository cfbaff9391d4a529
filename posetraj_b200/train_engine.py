"""One whole training step of BASELINE configs[3] on the sm_100a library: ControlNet forward, frozen UNet forward, EDM loss,
the "spatial" auxiliary pass, the reverse pass through both networks into the ControlNet's parameters, data-parallel
gradient averaging and AdamW (SURVEY.md 8e "training DP", 8f row 4).

Reference: /root/reference/scripts/train_svd_traj_VIPSeg_14_cam_concat.py
    :1320-1361  noise / sigma / conditioning latents / timesteps / added_time_ids        (ControlNetTrainer._stage)
    :1404-1412  ControlNet, then the UNet with its residuals                              (NetPlan op lists, train=True)
    :1417-1436  EDM-weighted MSE                                                          (pt_edm_loss)
    :1438-1462  second UNet pass on ONE frame with the residuals sliced `sample[ran_idx]` (spatial plan, weight 0.5)
    :1470       accelerator.backward(loss): autograd through UNet and ControlNet          (Tape: this file)
    :1165,1472  DDP gradient averaging + AdamW                                            (training.GradientBuckets / AdamW)

How the reverse pass is built.  posetraj_b200.engine.NetPlan lowers a network to a flat list of kernel launches; every
launch descriptor keeps a semantic record (`op.io`).  `Tape` walks such a list BACKWARDS and, per launch, runs the
backward kernels of that operator (dgrad = pt_gemm with W^T and negated taps, wgrad = pt_wgrad, GroupNorm / LayerNorm /
GEGLU / SiLU / attention / up-down-sampling backward, grouped column sums for bias, time-embedding and cross-attention
constants, dot products for the AlphaBlender mix factors).  Gradients of activations are bf16 tensors keyed by the
forward buffer; a forward sweep first marks which buffers depend on a trainable parameter, so the frozen UNet is only
differentiated from the loss back to the 13 injected residuals (its encoder is never touched).  There is no autograd
and no fallback anywhere: these functions are the product path and raise without the CUDA library.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib, ops, training
from .config import SVDConfig, controlnet_param_shapes
from .engine import BF16, F32, NetPlan, WeightStore


def _sp() -> int:
    return torch.cuda.current_stream().cuda_stream


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _key(t: torch.Tensor) -> int:
    return t.untyped_storage().data_ptr()


# ===================================================================================================================
# which tensors an op reads / writes (for the requires-grad sweep)
# ===================================================================================================================
def _op_tensors(op):
    """(inputs, outputs, has_parameters) of a launch descriptor, or None for ops that only stage inputs."""
    io = getattr(op, "io", None)
    if io is None:
        return None
    if isinstance(op, ops.Gemm):
        return [io.a0, io.a1, io.res1, io.res2, io.aux, io.rowvec], [io.out, io.out2], True
    if isinstance(op, ops.GroupNorm):
        return [io.x0, io.x1], [io.out], True
    if isinstance(op, ops.LayerNorm):
        return [io.x, io.addvec], [io.out, io.sum_out], True
    if isinstance(op, (ops.AttnSpatial, ops.AttnTemporal)):
        return [io.qkv], [io.out], False
    if isinstance(op, ops.SmallLinear):
        return [io.x], [io.out], True
    if isinstance(op, ops.Upsample2x):
        return [io.x], [io.out], False
    if isinstance(op, ops.Axpy):
        return [io.x, io.y], [io.out], False
    if isinstance(op, ops.GegluFwd):
        return [io.h], [io.out], False
    if isinstance(op, ops.SiluFwd):
        return [io.x], [io.out], False
    return None


class Tape:
    """Reverse-mode pass over NetPlan op lists (see the module docstring)."""

    def __init__(self, plan: NetPlan, *, trainable: bool, on_weight_done: Optional[Callable] = None):
        self.plan, self.trainable, self.w = plan, trainable, plan.w
        self.g: Dict[int, list] = {}          # storage ptr of a bf16 forward buffer -> [gradient, owned]
        self.gs: Dict[int, torch.Tensor] = {}  # storage ptr of a small fp32 forward buffer -> zero-initialised gradient
        self.wg: Dict[tuple, torch.Tensor] = {}  # (kind, key) of the WeightStore -> fp32 gradient in the KERNEL layout
        self.need: set = set()
        self.on_weight_done = on_weight_done
        self._uses: Dict[tuple, int] = {}
        self.flushed: set = set()
        # every Tape of a trained plan is a new optimizer step: cached transposed weights are stale
        self.step_id = plan.__dict__["_tape_steps"] = plan.__dict__.get("_tape_steps", 0) + 1

    # ---------------------------------------------------------------------------------------------------------
    # requires-grad sweep
    # ---------------------------------------------------------------------------------------------------------
    def mark(self, op_lists: Sequence[List], seeds: Sequence[torch.Tensor] = ()) -> None:
        for t in seeds:
            self.need.add(_key(t))
        self._uses = {}
        for lst in op_lists:
            for op in lst:
                tio = _op_tensors(op)
                if tio is None:
                    continue
                ins, outs, has_params = tio
                if isinstance(op, ops.Gemm) and op.io.out2 is not None:
                    # out2 = out + aux_scale * aux: `out` itself does not depend on aux (the frozen UNet's encoder never
                    # depends on the injected residuals, so its reverse pass is never run)
                    io = op.io
                    main = (self.trainable and has_params) or any(t is not None and _key(t) in self.need
                                                                  for t in (io.a0, io.a1, io.res1, io.res2, io.rowvec))
                    if main:
                        self.need.add(_key(io.out))
                    if main or _key(io.aux) in self.need:
                        self.need.add(_key(io.out2))
                elif (self.trainable and has_params) or any(t is not None and _key(t) in self.need for t in ins):
                    for t in outs:
                        if t is not None:
                            self.need.add(_key(t))
                if self.trainable and has_params:
                    for wt in self._op_weights(op):
                        k = self.w.origin(wt)
                        self._uses[k] = self._uses.get(k, 0) + 1

    @staticmethod
    def _op_weights(op):
        io = op.io
        if isinstance(op, ops.Gemm):
            return [t for t in (io.w, io.bias) if t is not None]
        if isinstance(op, (ops.GroupNorm, ops.LayerNorm)):
            return [io.gamma, io.beta]
        if isinstance(op, ops.SmallLinear):
            return [t for t in (io.w, io.bias) if t is not None]
        return []

    def needs(self, t: Optional[torch.Tensor]) -> bool:
        return t is not None and _key(t) in self.need

    # ---------------------------------------------------------------------------------------------------------
    # activation gradients (bf16)
    # ---------------------------------------------------------------------------------------------------------
    def seed(self, t: torch.Tensor, grad: torch.Tensor) -> None:
        self.need.add(_key(t))
        self.g[_key(t)] = [grad, False]

    def pop(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is None:
            return None
        e = self.g.pop(_key(t), None)
        return None if e is None else e[0]

    def add(self, t: Optional[torch.Tensor], c: torch.Tensor, scale: float = 1.0) -> None:
        """grad(t) += scale * c.  A first contribution with scale 1 is kept by reference (it may be shared with other
        gradients, so it is never written in place: `owned` is False until a sum has been materialised)."""
        if t is None or _key(t) not in self.need:
            return
        k = _key(t)
        cur = self.g.get(k)
        sp = _sp()
        if cur is None:
            if scale == 1.0:
                self.g[k] = [c, False]
            else:
                out = torch.empty(c.shape, device=c.device, dtype=BF16)
                ops.Axpy(c, c, out, scale - 1.0).launch(sp)      # c + (scale - 1) c
                self.g[k] = [out, True]
            return
        tgt = cur[0] if cur[1] else torch.empty(cur[0].shape, device=c.device, dtype=BF16)
        ops.Axpy(cur[0], c, tgt, scale).launch(sp)
        self.g[k] = [tgt, True]

    def gsmall(self, t: torch.Tensor) -> torch.Tensor:
        """fp32 gradient VIEW matching a small fp32 forward tensor (time embeddings, cross-attention constants, ...):
        one zero-initialised buffer per storage, every contribution accumulates."""
        k = _key(t)
        base = self.gs.get(k)
        if base is None:
            base = torch.zeros(t.untyped_storage().nbytes() // 4, device=t.device, dtype=F32)
            self.gs[k] = base
        return torch.as_strided(base, t.shape, t.stride(), t.storage_offset())

    # ---------------------------------------------------------------------------------------------------------
    # weight gradients (fp32, kernel layout)
    # ---------------------------------------------------------------------------------------------------------
    def _weight_used(self, wt: torch.Tensor) -> None:
        k = self.w.origin(wt)
        self._uses[k] -= 1
        if self._uses[k] == 0 and self.on_weight_done is not None:
            self.flushed.add(k)
            self.on_weight_done(k, self.wg.get(k))

    def flush_rest(self) -> None:
        """Weights whose use count never reached zero (an op that received no gradient) and the mix factors."""
        if self.on_weight_done is None:
            return
        for k, g in list(self.wg.items()):
            if k not in self.flushed:
                self.flushed.add(k)
                self.on_weight_done(k, g)

    def acc_vec(self, wt: torch.Tensor, v: torch.Tensor) -> None:
        k = self.w.origin(wt)
        if k in self.wg:
            self.wg[k].add_(v)          # a handful of floats (norm scales / shifts)
        else:
            self.wg[k] = v
        self._weight_used(wt)

    def acc_colsum(self, wt: torch.Tensor, x: torch.Tensor, scale: float) -> None:
        k = self.w.origin(wt)
        cur = self.wg.get(k)
        n = x.shape[1]
        if cur is None:
            cur = self.wg[k] = torch.zeros(wt.numel(), device=x.device, dtype=F32)
        training.colsum_grouped(x, groups=1, mode=1, ga=x.shape[0], scale=scale, out=cur[:n].view(1, n), accumulate=True)
        self._weight_used(wt)

    def dgrad_weight(self, w: torch.Tensor, T: int, n_pad: int) -> torch.Tensor:
        """[N, T*K] -> [K, T*n_pad] with Wd[k, t*n_pad + n] = W[n, t*K + k] (zero columns for padded output channels):
        the B operand of the dgrad GEMM.  Kept on the plan: computed once for frozen weights, re-derived in place (one
        pt_transpose_bf16 per tap) the first time it is needed in every step for trained ones."""
        cache = self.plan.__dict__.setdefault("_dgrad_w", {})
        k = (w.data_ptr(), T, n_pad)
        ent = cache.get(k)
        if ent is not None and (not self.trainable or ent[1] == self.step_id):
            return ent[0]
        N, K = w.shape[0], w.shape[1] // T
        wd = ent[0] if ent is not None else torch.zeros(K, T * n_pad, device=w.device, dtype=BF16)
        sp = _sp()
        for t in range(T):
            _lib.check(_lib.lib().pt_transpose_bf16(w.data_ptr() + 2 * t * K, w.stride(0), wd.data_ptr() + 2 * t * n_pad, wd.stride(0), N, K, sp),
                       "pt_transpose_bf16")
        cache[k] = (wd, self.step_id)
        return wd

    # ---------------------------------------------------------------------------------------------------------
    # driver
    # ---------------------------------------------------------------------------------------------------------
    def backward(self, op_lists: Sequence[List]) -> None:
        for lst in reversed(list(op_lists)):
            for op in reversed(lst):
                rule = _RULES.get(type(op))
                if rule is not None:
                    rule(self, op)


# ===================================================================================================================
# per-operator rules
# ===================================================================================================================
def _bw_gemm(T: Tape, op: ops.Gemm) -> None:
    io = op.io
    g1 = T.pop(io.out)
    g2 = T.pop(io.out2) if io.out2 is not None else None
    if g2 is not None:
        T.add(io.aux, g2, io.aux_scale)                     # out2 = val + aux_scale * aux
    if g1 is None and g2 is None:
        return
    if io.geglu or io.act or io.scatter is not None:
        raise RuntimeError(f"{op.name}: fused GEGLU / activation / scatter epilogues are not differentiated (train=True plans do not emit them)")
    if g1 is not None and g2 is not None:
        dval = torch.empty(g1.shape, device=g1.device, dtype=BF16)
        ops.Axpy(g1, g2, dval, 1.0).launch(_sp())
    else:
        dval = g1 if g1 is not None else g2
    N = io.n_out
    dv = dval if dval.shape[1] == N else dval[:, :N]
    # AlphaBlender folded into this epilogue: out = alpha * P + (1 - alpha) * Q, so sum dval o (P - Q) =
    # (dot(dval, P) - dot(dval, out)) / (1 - alpha) and d mix_factor = alpha (1 - alpha) * that
    if T.trainable and io.mix is not None:
        key, P, alpha = io.mix
        acc = T.wg.get(("mix", key))
        if acc is None:
            acc = T.wg[("mix", key)] = torch.zeros(1, device=dval.device, dtype=F32)
        ws = torch.empty(8192, device=dval.device, dtype=torch.uint8)
        for other, sgn in ((P, alpha), (io.out, -alpha)):
            _lib.check(_lib.lib().pt_dot_bf16(dv.data_ptr(), dv.stride(0), other.data_ptr(), other.stride(0), dv.shape[0], N, sgn,
                                              acc.data_ptr(), 1, ws.data_ptr(), _sp()), "pt_dot_bf16")
    for res, s in ((io.res1, io.res1_scale), (io.res2, io.res2_scale)):
        if res is None:
            continue
        if res.data_ptr() == io.out.data_ptr():               # in-place accumulation (conv_out of the second tower)
            assert s == 1.0
            T.g[_key(io.out)] = [dval, False]
        else:
            T.add(res, dval, s)
    scale = io.acc_scale * (T.plan._cond_scale_host if io.acc_scale_dev is not None else 1.0)
    need_a = T.needs(io.a0) or T.needs(io.a1)
    need_w = T.trainable
    need_rv = io.rowvec is not None and T.needs(io.rowvec)
    if need_w and io.bias is not None:
        T.acc_colsum(io.bias, dv, scale)
    if need_rv:
        gv = T.gsmall(io.rowvec)
        groups = io.rowvec.shape[0] if io.rowvec.dim() == 2 else 1
        a_, b_, c_ = io.rv[:3]
        training.colsum_grouped(dv, groups=groups, mode=io.rowvec_mode, ga=a_, gb=b_, gc=c_, scale=scale, out=gv[:, :N], accumulate=True)
    if not (need_a or need_w):
        return
    # the gradient in the row space of the A operand
    n_img = H = W = None
    if io.halo is None:
        dsp = dval
    else:
        H, W = io.halo
        n_img = io.a0.shape[0] // ((H + 1) * (W + 1))
        if dval.shape[1] % 8:
            dsp = torch.zeros(dval.shape[0], _pad64(dval.shape[1]), device=dval.device, dtype=BF16)
            dsp[:, : dval.shape[1]] = dval
            dval = dsp
        if io.ostride == 1:
            dsp = dval if io.out_halo else training.to_halo(dval, n_img, H, W)
        else:
            dsp = training.dilate2x(dval, n=n_img, H=H, W=W, src_halo=io.out_halo)
    if dsp.shape[1] % 64:
        buf = torch.zeros(dsp.shape[0], _pad64(dsp.shape[1]), device=dsp.device, dtype=BF16)
        buf[:, : dsp.shape[1]] = dsp
        dsp = buf
    n_pad = dsp.shape[1]
    if need_a:
        wd = T.dgrad_weight(io.w, len(io.taps), n_pad)
        dA = training.linear_dgrad(dsp, None, taps=io.taps, batches=io.batches, scale=scale, wd=wd)
        if io.halo is not None:
            training.zero_halo(dA, n=n_img, H=H, W=W)
        k0 = io.a0.shape[1]
        if io.a1 is None:
            T.add(io.a0, dA)
        else:
            T.add(io.a0, dA[:, :k0])
            T.add(io.a1, dA[:, k0:])
    if need_w:
        if io.a1 is not None:
            raise RuntimeError(f"{op.name}: weight gradient of a two-source (concat) GEMM is not implemented (frozen UNet only)")
        k = T.w.origin(io.w)
        cur = T.wg.get(k)
        T.wg[k] = training.wgrad(dsp, io.a0, taps=io.taps, scale=scale, batches=io.batches, out=cur, accumulate=cur is not None)
        T._weight_used(io.w)


def _bw_groupnorm(T: Tape, op: ops.GroupNorm) -> None:
    io = op.io
    g = T.pop(io.out)
    if g is None:
        return
    assert io.mode == 0
    dx0, dx1, dgb = training.groupnorm_backward(io.x0, g, io.gamma, io.beta, rows_per_stat=io.rows_per_stat, eps=io.eps, silu=io.silu,
                                                x1=io.x1, halo=io.halo, want_param_grads=T.trainable)
    T.add(io.x0, dx0)
    if io.x1 is not None:
        T.add(io.x1, dx1)
    if T.trainable:
        T.acc_vec(io.gamma, dgb[0])
        T.acc_vec(io.beta, dgb[1])


def _bw_layernorm(T: Tape, op: ops.LayerNorm) -> None:
    io = op.io
    g = T.pop(io.out)
    gs = T.pop(io.sum_out) if io.sum_out is not None else None
    if g is None and gs is None:
        return
    parts = []
    if g is not None:
        dx, dgb = training.layernorm_backward(io.x, g, io.gamma, eps=io.eps, want_param_grads=T.trainable, addvec=io.addvec, hw=io.hw,
                                              frames=io.frames)
        parts.append(dx)
        if T.trainable:
            T.acc_vec(io.gamma, dgb[0])
            T.acc_vec(io.beta, dgb[1])
    if gs is not None:
        parts.append(gs)
    for pt_ in parts:
        T.add(io.x, pt_)
        if io.addvec is not None and T.needs(io.addvec):   # d addvec[f] = sum over the rows of frame f
            training.colsum_grouped(pt_, groups=io.frames, mode=3, ga=io.hw, gc=io.frames, out=T.gsmall(io.addvec), accumulate=True)


def _bw_attn_spatial(T: Tape, op: ops.AttnSpatial) -> None:
    io = op.io
    g = T.pop(io.out)
    if g is None:
        return
    if io.lse is None:
        raise RuntimeError("spatial attention backward needs the forward log-sum-exp (build the plan with train=True)")
    T.add(io.qkv, training.attention_spatial_backward(io.qkv, io.out, g, io.lse, n_img=io.n_img, heads=io.heads))


def _bw_attn_temporal(T: Tape, op: ops.AttnTemporal) -> None:
    io = op.io
    g = T.pop(io.out)
    if g is None:
        return
    T.add(io.qkv, training.attention_temporal_backward(io.qkv, g, batch=io.batch, frames=io.frames, hw=io.hw, heads=io.heads))


def _bw_geglu(T: Tape, op: ops.GegluFwd) -> None:
    g = T.pop(op.io.out)
    if g is not None:
        T.add(op.io.h, training.geglu_backward(op.io.h, g))


def _bw_silu(T: Tape, op: ops.SiluFwd) -> None:
    g = T.pop(op.io.out)
    if g is not None:
        T.add(op.io.x, training.silu_backward(op.io.x, g))


def _bw_upsample(T: Tape, op: ops.Upsample2x) -> None:
    io = op.io
    g = T.pop(io.out)
    if g is not None:
        T.add(io.x, training.upsample_backward(g, n=io.n, H=io.H, W=io.W, halo=io.halo, scale=io.scale))


def _bw_axpy(T: Tape, op: ops.Axpy) -> None:
    io = op.io
    g = T.pop(io.out)
    if g is not None:
        T.add(io.x, g)
        T.add(io.y, g, io.scale)


def _bw_small_linear(T: Tape, op: ops.SmallLinear) -> None:
    io = op.io
    if _key(io.out) not in T.gs:
        return                                              # nothing flowed into this output
    if io.act_out_silu:
        raise RuntimeError(f"{op.name}: fused output SiLU is not differentiated (train=True plans move it to the consumer)")
    dy = T.gsmall(io.out)
    want_dx = T.needs(io.x)
    dw = db = None
    if T.trainable:
        kw, = [T.w.origin(io.w)]
        dw = T.wg.get(kw)
        first = dw is None
        if first:
            dw = T.wg[kw] = torch.zeros(io.w.shape, device=io.w.device, dtype=F32)
        if io.bias is not None:
            kb = T.w.origin(io.bias)
            db = T.wg.get(kb)
            if db is None:
                db = T.wg[kb] = torch.zeros(io.bias.numel(), device=io.w.device, dtype=F32)
    if not (want_dx or T.trainable):
        return
    a = _lib.PtSmallLinearBwdArgs()
    a.x, a.x_ld, a.w, a.w_ld, a.dy, a.dy_ld = io.x.data_ptr(), io.x.stride(0), io.w.data_ptr(), io.w.stride(0), dy.data_ptr(), dy.stride(0)
    a.M, a.K = io.x.shape
    a.N, a.act_in_silu = io.w.shape[0], int(io.act_in_silu)
    if want_dx:
        dx = T.gsmall(io.x)
        ws = torch.empty(_lib.lib().pt_small_linear_bwd_workspace_bytes(a.M, a.N, a.K), device=dx.device, dtype=torch.uint8)
        a.dx, a.dx_ld, a.accumulate_dx, a.dx_workspace = dx.data_ptr(), dx.stride(0), 1, ws.data_ptr()
    if T.trainable:
        a.dw, a.db, a.accumulate_w = dw.data_ptr(), (db.data_ptr() if db is not None else None), 1
    _lib.check(_lib.lib().pt_small_linear_bwd(training.C.addressof(a), _sp()), "pt_small_linear_bwd")
    if T.trainable:
        T._weight_used(io.w)
        if io.bias is not None:
            T._weight_used(io.bias)


_RULES = {ops.Gemm: _bw_gemm, ops.GroupNorm: _bw_groupnorm, ops.LayerNorm: _bw_layernorm, ops.AttnSpatial: _bw_attn_spatial,
          ops.AttnTemporal: _bw_attn_temporal, ops.GegluFwd: _bw_geglu, ops.SiluFwd: _bw_silu, ops.Upsample2x: _bw_upsample,
          ops.Axpy: _bw_axpy, ops.SmallLinear: _bw_small_linear}


# ===================================================================================================================
# kernel-layout weight gradients -> parameter-shaped gradients
# ===================================================================================================================
def to_param_grads(kind_key: tuple, g: Optional[torch.Tensor], shapes: Dict[str, tuple]) -> Dict[str, torch.Tensor]:
    """Inverse of the WeightStore re-layouts (engine.WeightStore.linear / conv3 / tconv / qkv / cat_rows / f32), applied to
    a gradient: {parameter name: fp32 gradient of the parameter's own shape}.  Pure re-layout (views + one copy)."""
    kind, key = kind_key
    if g is None:
        return {}
    if kind == "mix":
        a = g                                               # already d mix_factor
        return {key: a.reshape(shapes[key])}
    if kind == "f32":
        return {key: g[: math.prod(shapes[key])].reshape(shapes[key])}
    if isinstance(kind, tuple) and kind[0] in ("ccf", "ccc"):
        # one column block of cc_projection.weight: (gradient, column slice of the parameter)
        c_feat = kind[1]
        co = shapes[key][0]
        sl = (slice(None), slice(0, c_feat)) if kind[0] == "ccf" else (slice(None), slice(c_feat, shapes[key][1]))
        return {key: (g[:co, : (c_feat if kind[0] == "ccf" else shapes[key][1] - c_feat)], sl)}
    if kind == "lin":
        shp = shapes[key]
        return {key: g[: shp[0]].reshape(shp)}
    if isinstance(kind, tuple) and kind[0] == "conv3":
        co, ci = shapes[key][:2]
        cp = kind[1] or ci
        return {key: g[:co].view(co, 3, 3, cp)[..., :ci].permute(0, 3, 1, 2).contiguous()}
    if kind == "tconv":
        co, ci = shapes[key][:2]
        return {key: g[:co].view(co, 3, ci).permute(0, 2, 1).reshape(co, ci, 3, 1, 1).contiguous()}
    if kind == "qkv":
        out, r = {}, 0
        for n in ("q", "k", "v"):
            k = key + f"to_{n}.weight"
            out[k] = g[r: r + shapes[k][0]].reshape(shapes[k])
            r += shapes[k][0]
        return out
    if isinstance(kind, tuple) and kind[0] == "cat":
        out, r = {}, 0
        flat = kind[1] == "f32"
        for k in key:
            n0 = shapes[k][0]
            out[k] = (g.reshape(-1)[r: r + n0] if flat else g[r: r + n0]).reshape(shapes[k])
            r += n0
        return out
    raise KeyError(f"no gradient re-layout for weight kind {kind!r}")


# ===================================================================================================================
# the training step
# ===================================================================================================================
class ControlNetTrainer:
    """ControlNet fine-tuning step of the reference's training loop (see the module docstring) for a fixed problem shape.

    `unet` / `controlnet` are the posetraj_b200.models mirrors (the UNet stays frozen: train...cam_concat.py:984-987,1101);
    the ControlNet's parameters are copied into fp32 master buffers owned by the optimizer, and the bf16 kernel-ready
    copies are re-derived from them after every step.  `group`: torch.distributed process group of the data-parallel
    ranks (None: the default group when initialised, else single process)."""

    def __init__(self, unet, controlnet, *, batch: int, frames: int, height: int, width: int, lr: float = 1e-5, betas=(0.9, 0.999),
                 weight_decay: float = 1e-2, eps: float = 1e-8, use_spatial: bool = True, group=None, bucket_mb: float = 100.0):
        dev = controlnet.device
        self.dev, self.B, self.F, self.H, self.W = dev, batch, frames, height, width
        self.cfg: SVDConfig = controlnet.cfg
        self.bbox = bool(controlnet.flags.get("bbox"))
        self.cam = bool(controlnet.flags.get("cam"))
        self.shapes = controlnet_param_shapes(self.cfg, cam=self.cam, bbox=self.bbox)
        self.names = list(self.shapes.keys())
        self.index = {k: i for i, k in enumerate(self.names)}
        sizes = [int(math.prod(self.shapes[k])) for k in self.names]
        sd = controlnet.state_dict()
        self.buckets = training.GradientBuckets(sizes, dev, group=group, bucket_mb=bucket_mb)
        self.opt = training.AdamW(self.buckets, [sd[k] for k in self.names], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        # the state dict the kernels' weight store derives from: VIEWS of the optimizer's fp32 master buffers
        self.master = {k: self.opt.param(i).view(self.shapes[k]) for i, k in enumerate(self.names)}
        self.cw = WeightStore(self.master, dev)
        self.cplan = NetPlan("controlnet", self.cfg, self.cw, batch=batch, frames=frames, height=height, width=width, device=dev,
                             bbox=self.bbox, cam=self.cam, train=True)
        self.uplan = NetPlan("unet", unet.cfg, unet.weights, batch=batch, frames=frames, height=height, width=width, device=dev,
                             residual_bufs=self.cplan.res, x_in=self.cplan.x_in, train=True)
        self.use_spatial = use_spatial
        if use_spatial:
            self.splan = NetPlan("unet", unet.cfg, unet.weights, batch=batch, frames=1, height=height, width=width, device=dev, train=True)
        self.loss = torch.zeros(1, device=dev, dtype=F32)
        self._written: set = set()
        # CUDA-graph replay of forward + backward (one graph per `ran_idx`, sharing one memory pool): the eager step is
        # ~5 700 launches whose host cost (descriptor objects, ctypes calls, allocator) is as long as the device time
        self.use_cuda_graph = False
        self._graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self._pool = None
        self._static: Optional[Dict[str, torch.Tensor]] = None
        self.graph_launches = 0

    # ---------------------------------------------------------------------------------------------------------
    def _stage(self, plan: NetPlan, inp: torch.Tensor, timesteps, ehs, ids) -> None:
        sp = _sp()
        b, Fr = inp.shape[:2]
        ops.Layout(inp.reshape(b * Fr, *inp.shape[2:]).contiguous(), plan.x_in, to_tokens=True, halo=True).launch(sp)
        plan.t_buf.copy_(timesteps)
        plan.time_ids.copy_(ids.reshape(-1))
        plan.ehs.copy_(ehs[:, 0, :])
        NetPlan.run(plan.embed_ops, sp)

    def _refresh_alphas(self) -> None:
        ups = self.cplan.alpha_updaters
        if not ups:
            return
        vals = torch.sigmoid(torch.stack([self.master[k].reshape(-1)[0] for k, _ in ups])).cpu().tolist()
        for (_, fn), a in zip(ups, vals):
            fn(float(a))

    def _flush_weight(self, kind_key, g) -> None:
        """A weight's gradient is final: re-layout into the parameter's bucket slot and mark it ready (its bucket's
        all-reduce starts as soon as the bucket is complete — overlapped with the rest of the reverse pass)."""
        for name, pg in to_param_grads(kind_key, g, self.shapes).items():
            i = self.index[name]
            view = self.buckets.view(i).view(self.shapes[name])
            if isinstance(pg, tuple):       # a column block of a parameter that two launches share (cc_projection.weight)
                pg, sl = pg
                if i not in self._partial:
                    view.zero_()
                    self._partial[i] = 0
                view[sl].copy_(pg)
                self._partial[i] += 1
                if self._partial[i] < 2:
                    continue
            else:
                view.copy_(pg)
            self._written.add(i)
            self.buckets.ready(i)

    # ---------------------------------------------------------------------------------------------------------
    def forward_backward(self, *, latents, noise, sigmas, image_embeddings, trajectories, motion_values, controlnet_bbox=None,
                         camera_cond=None, ran_idx: int = 0, scaling_factor: float = 0.18215, noise_aug: float = 0.02,
                         random_p: Optional[torch.Tensor] = None, conditioning_dropout_prob: Optional[float] = None) -> torch.Tensor:
        """Loss and ControlNet gradients (into the gradient buckets; all-reduces launched).  Inputs as oracle/train.py
        `training_step`: latents [b, F, 4, h, w] (already x scaling_factor), noise like latents, sigmas [b],
        image_embeddings [b, 1, D], trajectories [b, F, 3, 8h, 8w], motion_values [b]; `camera_cond` [b, F, 12] for the
        camera model (scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1404-1412), `controlnet_bbox` like trajectories;
        `random_p` [b] in [0, 1) with `conditioning_dropout_prob`: the reference's conditioning dropout (:1365-1385 — the image
        embedding is zeroed where random_p < 2p, the conditioning latent where p <= random_p < 3p)."""
        sp = _sp()
        b, Fr = latents.shape[:2]
        assert (b, Fr) == (self.B, self.F)
        dev = self.dev
        latents, noise = latents.to(dev, F32), noise.to(dev, F32)
        sigmas = sigmas.to(dev, F32).reshape(b)
        # ---- input staging (train...cam_concat.py:1336-1361): a few elementwise ops on [b, F, 4, h, w]
        s = sigmas.view(b, 1, 1, 1, 1)
        cond = (latents + noise * noise_aug)[:, 0] / scaling_factor
        noisy = (latents + noise * s).contiguous()
        timesteps = 0.25 * sigmas.log()
        inp = torch.cat([noisy / ((s ** 2 + 1) ** 0.5), cond.unsqueeze(1).repeat(1, Fr, 1, 1, 1)], dim=2)
        ids = torch.stack([torch.full((b,), 6.0, device=dev), torch.full((b,), noise_aug, device=dev), motion_values.to(dev, F32).reshape(b)], 1)
        ehs = image_embeddings.to(dev, F32)
        if conditioning_dropout_prob is not None:
            if random_p is None:
                raise ValueError("conditioning_dropout_prob needs random_p (one uniform draw per sample)")
            rp = random_p.to(dev, F32).reshape(b)
            pd = float(conditioning_dropout_prob)
            ehs = torch.where((rp < 2 * pd).reshape(b, 1, 1), torch.zeros_like(ehs), ehs)
            image_mask = 1 - ((rp >= pd).to(F32) * (rp < 3 * pd).to(F32))
            cond = image_mask.reshape(b, 1, 1, 1) * cond
            inp = torch.cat([noisy / ((s ** 2 + 1) ** 0.5), cond.unsqueeze(1).repeat(1, Fr, 1, 1, 1)], dim=2)
        cp, up = self.cplan, self.uplan
        # ---- ControlNet
        self._stage(cp, inp, timesteps, ehs, ids)
        cp.cond_in.copy_(trajectories.to(dev, F32).reshape(cp.cond_in.shape))
        use_bbox = self.bbox and controlnet_bbox is not None
        if use_bbox:
            cp.cond_in2.copy_(controlnet_bbox.to(dev, F32).reshape(cp.cond_in2.shape))
        use_cam = self.cam and camera_cond is not None
        if camera_cond is not None and not self.cam:
            raise ValueError("camera_cond given but this ControlNet was built without cc_projection (cam=False)")
        if use_cam:
            cp.cam_in.copy_(camera_cond.to(dev, F32).reshape(cp.n, 12))
        cond_ops = cp.cond_op_list(use_cam, use_bbox)
        NetPlan.run(cond_ops, sp)
        NetPlan.run(cp.step_ops, sp)
        # ---- UNet on the same input (x_in is shared) with the residuals read in place
        up.t_buf.copy_(timesteps)
        up.time_ids.copy_(ids.reshape(-1))
        up.ehs.copy_(ehs[:, 0, :])
        NetPlan.run(up.embed_ops, sp)
        NetPlan.run(up.step_ops, sp)
        rows = up.noise_pred.shape[0]
        dpred = torch.zeros(rows, 64, device=dev, dtype=BF16)
        latc = latents.contiguous()
        training.edm_loss_into(up.noise_pred, noisy, latc, sigmas, weight=1.0, frame=None, loss=self.loss, accumulate=False,
                               dpred=dpred[:, : self.cfg.out_channels])
        # ---- "spatial" pass: ONE frame through the UNet, residuals sliced on the flattened (b*F) axis (:1438-1462)
        if self.use_spatial:
            spn = self.splan
            self._stage(spn, inp[:, ran_idx: ran_idx + 1], timesteps, ehs, ids)
            for i, (src, dst) in enumerate(zip(cp.res, spn.res)):
                hw_i = dst.shape[0] // b
                for bb in range(b):     # sample[ran_idx].unsqueeze(0) broadcasts over the batch
                    dst[bb * hw_i: (bb + 1) * hw_i].copy_(src[ran_idx * hw_i: (ran_idx + 1) * hw_i])
            NetPlan.run(spn.step_ops, sp)
            dpred_s = torch.zeros(spn.noise_pred.shape[0], 64, device=dev, dtype=BF16)
            training.edm_loss_into(spn.noise_pred, noisy, latc, sigmas, weight=0.5, frame=ran_idx, loss=self.loss, accumulate=True,
                                   dpred=dpred_s[:, : self.cfg.out_channels])
        # ---- reverse pass: UNet(s) from the loss to the residuals, then the ControlNet
        self._written = set()
        self._partial: Dict[int, int] = {}
        tu = Tape(up, trainable=False)
        tu.mark([up.step_ops], seeds=cp.res)
        tu.seed(up.noise_pred, dpred)
        tu.backward([up.step_ops])
        tc = Tape(cp, trainable=True, on_weight_done=self._flush_weight)
        op_lists = [cond_ops, cp.embed_ops, cp.step_ops]
        tc.mark(op_lists)
        if self.use_spatial:
            ts = Tape(self.splan, trainable=False)
            ts.mark([self.splan.step_ops], seeds=self.splan.res)
            ts.seed(self.splan.noise_pred, dpred_s)
            ts.backward([self.splan.step_ops])
        for i, r in enumerate(cp.res):
            g = tu.pop(r)
            if self.use_spatial:
                gsp = ts.pop(self.splan.res[i])
                if gsp is not None:
                    hw_i = gsp.shape[0] // b
                    if g is None:
                        g = torch.zeros(r.shape, device=dev, dtype=BF16)
                    else:
                        g2 = torch.empty(r.shape, device=dev, dtype=BF16)
                        g2.copy_(g)
                        g = g2
                    sl = g[ran_idx * hw_i: (ran_idx + 1) * hw_i]
                    for bb in range(b):
                        ops.Axpy(sl, gsp[bb * hw_i: (bb + 1) * hw_i], sl, 1.0).launch(sp)
            if g is not None:
                tc.seed(r, g)
        tc.backward(op_lists)
        tc.flush_rest()
        # parameters no launch touched (dead cross-attention queries / keys, norm2, conv_out_2 ...): zero gradient
        for i in range(len(self.names)):
            if i not in self._written:
                if i not in self._partial:          # (a half-written split parameter keeps its written block, the rest is zero)
                    self.buckets.view(i).zero_()
                self.buckets.ready(i)
        return self.loss

    def gradients(self) -> Dict[str, torch.Tensor]:
        """{parameter name: fp32 gradient} views of the buckets (summed over ranks after `finish`)."""
        return {k: self.buckets.view(i).view(self.shapes[k]) for i, k in enumerate(self.names)}

    def optimizer_step(self) -> None:
        self.buckets.finish()
        self.opt.step()
        self.cw.refresh()
        self._refresh_alphas()

    def step(self, *, ran_idx: int = 0, **batch) -> torch.Tensor:
        """One optimizer step.  With `use_cuda_graph` the first call runs eagerly (it fills every cache: kernel attributes,
        the frozen UNet's transposed weights, NCCL communicators); from the second call on forward + backward (and the
        bucket all-reduces) replay as one captured graph per `ran_idx`, fed through static input buffers."""
        if not self.use_cuda_graph:
            loss = self.forward_backward(ran_idx=ran_idx, **batch)
            self.optimizer_step()
            return loss
        batch = {k: v for k, v in batch.items() if v is not None}
        if self._static is None:
            self._static = {k: v.detach().to(self.dev).clone() for k, v in batch.items()}
            loss = self.forward_backward(ran_idx=ran_idx, **self._static)
            self.optimizer_step()
            return loss
        if set(batch) != set(self._static):
            raise ValueError("cuda-graph training step: the set of inputs must not change between steps")
        for k, v in batch.items():
            self._static[k].copy_(v, non_blocking=True)
        g = self._graphs.get(ran_idx)
        if g is None:
            from . import _lib as _l
            n0 = _l.lib().pt_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool):
                self.forward_backward(ran_idx=ran_idx, **self._static)
                self.buckets.finish()           # joins the all-reduce stream back into the capture
            self.graph_launches = int(_l.lib().pt_launch_count() - n0)
            self._graphs[ran_idx] = g
            self._pool = g.pool()
        g.replay()
        self.optimizer_step()
        return self.loss

    def set_lr(self, lr: float) -> None:
        """`lr_scheduler.step()` of the reference loop (:1474): the learning rate is a launch argument of the AdamW kernel."""
        self.opt.lr = float(lr)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.clone() for k, v in self.master.items()}

    def save_pretrained(self, directory: str, variant: Optional[str] = None) -> str:
        """The trained ControlNet in the diffusers layout (`config.json` + `diffusion_pytorch_model.safetensors`, fp32) —
        what the reference writes with `controlnet.save_pretrained` at its checkpoints and what
        `ControlNetSDVModel.from_pretrained` (and the reference's own loader) read back."""
        from . import checkpoint
        return checkpoint.save_pretrained(self.state_dict(), self.cfg, directory, "ControlNetSDVModel", variant)
