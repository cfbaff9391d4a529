"""Training-step building blocks of BASELINE configs[3] on the sm_100a library (SURVEY.md 8f row 4, 8e "training DP").

Reference: /root/reference/scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1404-1475 — ControlNet forward, frozen UNet
forward, EDM-weighted MSE (:1423-1436), `accelerator.backward(loss)` (:1470) through both networks into the ControlNet's
parameters, AdamW (:1472); `accelerator.prepare` wraps the ControlNet in DDP (:1165), i.e. gradients are averaged over
the data-parallel ranks before the optimizer step.

What is here (each checked against torch autograd of the oracle on the GPU, tests/test_training_gpu.py):
  * the backward OPERATORS in the library's layouts — `linear_backward` / `conv_backward` (dgrad = pt_gemm with W^T and
    negated taps, wgrad = pt_wgrad on tcgen05, bias / time-embedding row-vector gradients = pt_colsum),
    `groupnorm_backward`, `layernorm_backward`, GEGLU forward / backward as a separate pass, `edm_loss`;
  * two block-level forward + backward passes assembled from them: `ResBlockTrainer` (SpatioTemporalResBlock: 4-D and 5-D
    GroupNorm+SiLU, 3x3 conv, temporal (3,1,1) conv, time-embedding injection, 1x1 shortcut, AlphaBlender incl. the
    mix_factor gradient) and `FeedForwardTrainer` (LayerNorm -> GEGLU -> Linear + residual);
  * the data-parallel step around them: `GradientBuckets` (flat fp32 buckets, one NCCL all-reduce per bucket launched as
    soon as the bucket's last gradient is written, i.e. overlapped with the rest of the backward pass; gloo on CPU for
    the tests) and `AdamW` (fused pt_adamw on fp32 master weights + the bf16 copies the kernels read).
What is NOT here yet: the attention backward (spatial flash backward, temporal) and the plan-level reverse pass that
would chain the blocks of both networks — see DESIGN.md "Training".  There is no autograd fallback: these functions are
the product path and raise without the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops

BF16, F32 = torch.bfloat16, torch.float32


def _sp() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check(t: torch.Tensor, dtype) -> None:
    if t.device.type != "cuda":
        raise RuntimeError("posetraj_b200.training runs on CUDA sm_100a only; there is no CPU path")
    assert t.dtype == dtype, (t.dtype, dtype)


# ---------------------------------------------------------------------------------------------------------------
# small kernels
# ---------------------------------------------------------------------------------------------------------------
def transpose(x: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, cols] -> [cols, rows] with the row stride padded to a multiple of 8 (TMA: 16-byte strides)."""
    _check(x, BF16)
    rows, cols = x.shape
    ld = (rows + 7) // 8 * 8
    buf = torch.zeros(cols, ld, device=x.device, dtype=BF16)
    for r0 in range(0, rows, 32 * 65535):           # grid.y limit
        n = min(rows - r0, 32 * 65535)
        _lib.check(_lib.lib().pt_transpose_bf16(x[r0:].data_ptr(), x.stride(0), buf[:, r0:].data_ptr(), ld, n, cols, _sp()),
                   "pt_transpose_bf16")
    return buf[:, :rows]


def colsum(x: torch.Tensor, *, groups: int = 1, halo: Optional[tuple] = None, rows: Optional[int] = None, scale: float = 1.0,
           out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """fp32 [groups, C] column sums of bf16 rows; `halo=(H, W)`: x is in the zero-haloed layout, `rows` logical pixels."""
    _check(x, BF16)
    Cc = x.shape[1]
    logical = rows if rows is not None else x.shape[0]
    assert logical % groups == 0
    if out is None:
        out = torch.zeros(groups, Cc, device=x.device, dtype=F32)
    a = _lib.PtColsumArgs()
    a.x, a.ld = x.data_ptr(), x.stride(0)
    if halo is not None:
        a.halo, a.H, a.W = 1, halo[0], halo[1]
    a.rows_per_group, a.groups, a.C, a.scale = logical // groups, groups, Cc, scale
    a.out, a.accumulate = out.data_ptr(), int(accumulate)
    _lib.check(_lib.lib().pt_colsum(C.addressof(a), _sp()), "pt_colsum")
    return out


def dot(a_: torch.Tensor, b_: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    _check(a_, BF16), _check(b_, BF16)
    assert a_.shape == b_.shape
    out = torch.zeros(1, device=a_.device, dtype=F32)
    ws = torch.empty(8192, device=a_.device, dtype=torch.uint8)
    _lib.check(_lib.lib().pt_dot_bf16(a_.data_ptr(), a_.stride(0), b_.data_ptr(), b_.stride(0), a_.shape[0], a_.shape[1], scale,
                                      out.data_ptr(), 0, ws.data_ptr(), _sp()), "pt_dot_bf16")
    return out


def to_halo(x: torch.Tensor, n: int, H: int, W: int) -> torch.Tensor:
    """compact [n*H*W, C] -> zero-haloed [n*(H+1)*(W+1), C] (the layout conv inputs / conv output gradients live in)."""
    out = torch.empty(n * (H + 1) * (W + 1), x.shape[1], device=x.device, dtype=BF16)   # the kernel writes the halo rows (zeros) too
    ops.Upsample2x(x, out, n=n, H=H, W=W, halo=True, scale=1).launch(_sp())
    return out


# ---------------------------------------------------------------------------------------------------------------
# linear / implicit-GEMM conv backward
# ---------------------------------------------------------------------------------------------------------------
def dgrad_weight(w: torch.Tensor, taps: int) -> torch.Tensor:
    """Forward weight [N, taps*K] (K index = t*K + k) -> dgrad weight [K, taps*N] with Wd[k, t*N + n] = W[n, t*K + k].
    (A pure re-layout of the parameters, done with a torch view/permute copy once per optimizer step.)"""
    N = w.shape[0]
    K = w.shape[1] // taps
    return w.view(N, taps, K).permute(2, 1, 0).reshape(K, taps * N).contiguous()


def linear_dgrad(dout: torch.Tensor, w: torch.Tensor, *, taps: Sequence[int] = (0,), batches: int = 1, scale: float = 1.0,
                 accumulate_into: Optional[torch.Tensor] = None, wd: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dA[r] = scale * sum_t dD[r - s_t] W_t  (+ accumulate_into): the forward kernel with negated shifts and W^T
    (oracle/backward.py conv_rows_dgrad).  `dout` lives in the SAME row space as the forward's A operand (zero-haloed for
    3x3 convs; rows that are not real outputs must be zero)."""
    _check(dout, BF16)
    T = len(taps)
    wd = wd if wd is not None else dgrad_weight(w, T)
    K = wd.shape[0]
    out = torch.empty(dout.shape[0], K, device=dout.device, dtype=BF16)
    ops.Gemm(dout, wd, out, taps=[-int(s) for s in taps], batches=batches, acc_scale=scale, res1=accumulate_into,
             name="dgrad").launch(_sp())
    return out


def wgrad(dout: torch.Tensor, a: torch.Tensor, *, taps: Sequence[int] = (0,), splits: Optional[int] = None,
          scale: float = 1.0, out: Optional[torch.Tensor] = None, accumulate: bool = False, batches: int = 1) -> torch.Tensor:
    """fp32 dW[n, t*K + k] = scale * sum_r dD[r, n] A[r + s_t, k] (pt_wgrad on tcgen05, both operands read as they are,
    + fixed-order fold of the row slices).  `dout` [rows, N] and `a` [rows, K] share one row space (zero-haloed for convs).  `batches` > 1: the taps
    must not reach across batch rows (temporal convs: frame f +- 1 of the SAME video), so each batch is its own launch
    whose out-of-range rows are zero-filled by TMA, exactly like the forward's rank-3 tensor map."""
    _check(dout, BF16), _check(a, BF16)
    if batches > 1:
        rpb = dout.shape[0] // batches
        for b in range(batches):
            out = wgrad(dout[b * rpb:(b + 1) * rpb], a[b * rpb:(b + 1) * rpb], taps=taps, splits=splits, scale=scale, out=out,
                        accumulate=accumulate or b > 0)
        return out
    rows, N = dout.shape
    K = a.shape[1]
    assert a.shape[0] == rows and K % 64 == 0 and a.stride(1) == 1
    T = len(taps)
    assert dout.stride(1) == 1 and dout.stride(0) % 8 == 0 and N % 8 == 0
    tiles = T * ((N + 127) // 128) * ((K // 64 + 3) // 4)
    if splits is None:
        splits = max(1, min((rows + 63) // 64, (2 * ops.NUM_SMS + tiles - 1) // tiles))
    partials = torch.empty(splits, N, T * K, device=dout.device, dtype=F32)
    tm_dt = _lib.encode_tensormap(dout.data_ptr(), [N, rows], [dout.stride(0) * 2], [64, 64])
    tm_a = _lib.encode_tensormap(a.data_ptr(), [K, rows], [a.stride(0) * 2], [64, 64])
    args = _lib.PtWgradArgs()
    args.tmap_dt, args.tmap_a = C.addressof(tm_dt), C.addressof(tm_a)
    args.rows, args.N, args.K, args.num_taps = rows, N, K, T
    for i, s in enumerate(taps):
        args.tap_shift[i] = int(s)
    args.splits, args.partials = splits, partials.data_ptr()
    _lib.check(_lib.lib().pt_wgrad(C.addressof(args), _sp()), "pt_wgrad")
    if out is None:
        out = torch.zeros(N, T * K, device=dout.device, dtype=F32)
    _lib.check(_lib.lib().pt_reduce_partials(partials.data_ptr(), splits, N * T * K, scale, out.data_ptr(), int(accumulate), _sp()),
               "pt_reduce_partials")
    return out


# ---------------------------------------------------------------------------------------------------------------
# normalisation / activation backward
# ---------------------------------------------------------------------------------------------------------------
def groupnorm_backward(x0: torch.Tensor, dout: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, rows_per_stat: int,
                       eps: float, silu: bool = True, x1: Optional[torch.Tensor] = None, halo: Optional[tuple] = None,
                       want_param_grads: bool = True):
    """GroupNorm(32)(+SiLU) backward (oracle/backward.py groupnorm_silu_backward).  Returns (dx0, dx1 | None, dgb | None)
    with dgb = fp32 [2, C] (dgamma, dbeta).  `dout` is in the forward output's layout (`halo=(H, W)`: zero-haloed)."""
    _check(x0, BF16), _check(dout, BF16)
    rows = x0.shape[0]
    c0, c1 = x0.shape[1], (x1.shape[1] if x1 is not None else 0)
    Cc = c0 + c1
    num_stat = rows // rows_per_stat
    dx0 = torch.empty(rows, c0, device=x0.device, dtype=BF16)
    dx1 = torch.empty(rows, c1, device=x0.device, dtype=BF16) if x1 is not None else None
    ws = torch.empty(_lib.lib().pt_groupnorm_bwd_workspace_bytes(num_stat, Cc), device=x0.device, dtype=torch.uint8)
    dgb = torch.zeros(2, Cc, device=x0.device, dtype=F32) if want_param_grads else None
    a = _lib.PtGroupNormBwdArgs()
    a.x0, a.c0, a.ld0 = x0.data_ptr(), c0, x0.stride(0)
    if x1 is not None:
        a.x1, a.c1, a.ld1 = x1.data_ptr(), c1, x1.stride(0)
        a.dx1, a.dld1 = dx1.data_ptr(), dx1.stride(0)
    a.dout, a.dout_ld = dout.data_ptr(), dout.stride(0)
    if halo is not None:
        a.halo, a.H, a.W = 1, halo[0], halo[1]
    a.gamma, a.beta, a.eps, a.silu = gamma.data_ptr(), beta.data_ptr(), eps, int(silu)
    a.rows_per_stat, a.num_stat = rows_per_stat, num_stat
    a.dx0, a.dld0 = dx0.data_ptr(), dx0.stride(0)
    a.workspace = ws.data_ptr()
    if dgb is not None:
        a.dgb_out = dgb.data_ptr()
    _lib.check(_lib.lib().pt_groupnorm_bwd(C.addressof(a), _sp()), "pt_groupnorm_bwd")
    return dx0, dx1, dgb


def layernorm_backward(x: torch.Tensor, dout: torch.Tensor, gamma: torch.Tensor, *, eps: float = 1e-5,
                       accumulate_into: Optional[torch.Tensor] = None, want_param_grads: bool = True,
                       addvec: Optional[torch.Tensor] = None, hw: int = 1, frames: int = 1):
    """LayerNorm backward (oracle/backward.py layernorm_backward).  Returns (dx, dgb | None); with `accumulate_into` the
    result is added to that tensor in place (the input also feeds a residual branch)."""
    _check(x, BF16), _check(dout, BF16)
    rows, Cc = x.shape
    dx = accumulate_into if accumulate_into is not None else torch.empty(rows, Cc, device=x.device, dtype=BF16)
    nb = min(ops.NUM_SMS * 2, (rows + 7) // 8)
    a = _lib.PtLayerNormBwdArgs()
    a.x, a.ld, a.dout, a.dout_ld = x.data_ptr(), x.stride(0), dout.data_ptr(), dout.stride(0)
    a.gamma, a.eps, a.rows, a.C = gamma.data_ptr(), eps, rows, Cc
    if addvec is not None:      # the forward normalised x + addvec[frame of the row]
        _check(addvec, F32)
        a.addvec, a.hw, a.F = addvec.data_ptr(), hw, frames
    a.dx, a.dx_ld, a.accumulate_dx = dx.data_ptr(), dx.stride(0), int(accumulate_into is not None)
    a.n_blocks = nb
    dgb = None
    if want_param_grads:
        partials = torch.empty(nb, 2 * Cc, device=x.device, dtype=F32)
        dgb = torch.zeros(2, Cc, device=x.device, dtype=F32)
        a.partials, a.dgb_out = partials.data_ptr(), dgb.data_ptr()
    _lib.check(_lib.lib().pt_layernorm_bwd(C.addressof(a), _sp()), "pt_layernorm_bwd")
    return dx, dgb


def geglu_forward(h: torch.Tensor) -> torch.Tensor:
    _check(h, BF16)
    rows, H2 = h.shape
    out = torch.empty(rows, H2 // 2, device=h.device, dtype=BF16)
    _lib.check(_lib.lib().pt_geglu_fwd(h.data_ptr(), h.stride(0), out.data_ptr(), out.stride(0), rows, H2 // 2, _sp()), "pt_geglu_fwd")
    return out


def geglu_backward(h: torch.Tensor, dout: torch.Tensor) -> torch.Tensor:
    _check(h, BF16), _check(dout, BF16)
    rows, H2 = h.shape
    dh = torch.empty_like(h)
    _lib.check(_lib.lib().pt_geglu_bwd(h.data_ptr(), h.stride(0), dout.data_ptr(), dout.stride(0), dh.data_ptr(), dh.stride(0), rows,
                                       H2 // 2, _sp()), "pt_geglu_bwd")
    return dh


def edm_loss(pred_tokens: torch.Tensor, noisy: torch.Tensor, target: torch.Tensor, sigmas: torch.Tensor, *, weight: float = 1.0,
             frame: Optional[int] = None, loss: Optional[torch.Tensor] = None, want_grad: bool = True):
    """EDM-weighted MSE of the reference (train...cam_concat.py:1417-1436) on the UNet's token-major prediction
    [B*F*HW, C]; `frame` selects one frame of noisy / target for the F = 1 "spatial" pass (:1438-1462, weight 0.5).
    Returns (loss fp32 [1] — accumulated into `loss` when given —, d loss / d pred as bf16 tokens)."""
    _check(pred_tokens, BF16), _check(noisy, F32), _check(target, F32), _check(sigmas, F32)
    B, Ft, Cc, H, W = noisy.shape
    Fr = 1 if frame is not None else Ft
    assert pred_tokens.shape[0] == B * Fr * H * W and noisy.is_contiguous() and target.is_contiguous()
    off = (frame or 0) * Cc * H * W
    dpred = torch.empty_like(pred_tokens) if want_grad else None
    acc = loss is not None
    if loss is None:
        loss = torch.zeros(1, device=noisy.device, dtype=F32)
    ws = torch.empty(_lib.lib().pt_edm_loss_workspace_bytes(), device=noisy.device, dtype=torch.uint8)
    a = _lib.PtEdmLossArgs()
    a.pred, a.pred_ld = pred_tokens.data_ptr(), pred_tokens.stride(0)
    a.noisy, a.target = noisy.data_ptr() + 4 * off, target.data_ptr() + 4 * off
    a.sample_stride, a.frame_stride = Ft * Cc * H * W, Cc * H * W
    a.sigmas, a.B, a.F, a.C, a.HW, a.weight = sigmas.data_ptr(), B, Fr, Cc, H * W, weight
    if dpred is not None:
        a.dpred, a.dpred_ld = dpred.data_ptr(), dpred.stride(0)
    a.workspace, a.loss, a.accumulate = ws.data_ptr(), loss.data_ptr(), int(acc)
    _lib.check(_lib.lib().pt_edm_loss(C.addressof(a), _sp()), "pt_edm_loss")
    return loss, dpred


def edm_loss_into(pred_tokens: torch.Tensor, noisy: torch.Tensor, target: torch.Tensor, sigmas: torch.Tensor, *, weight: float,
                  frame: Optional[int], loss: torch.Tensor, accumulate: bool, dpred: torch.Tensor) -> None:
    """`edm_loss` writing into caller-owned buffers: `loss` (+)= the weighted loss, `dpred` (a bf16 [rows, C] view, any row
    stride) = weight * d loss / d pred."""
    _check(pred_tokens, BF16), _check(noisy, F32), _check(target, F32), _check(sigmas, F32), _check(dpred, BF16)
    B, Ft, Cc, H, W = noisy.shape
    Fr = 1 if frame is not None else Ft
    assert pred_tokens.shape[0] == B * Fr * H * W and noisy.is_contiguous() and target.is_contiguous()
    assert dpred.shape[0] == pred_tokens.shape[0] and dpred.shape[1] == Cc
    off = (frame or 0) * Cc * H * W
    ws = torch.empty(_lib.lib().pt_edm_loss_workspace_bytes(), device=noisy.device, dtype=torch.uint8)
    a = _lib.PtEdmLossArgs()
    a.pred, a.pred_ld = pred_tokens.data_ptr(), pred_tokens.stride(0)
    a.noisy, a.target = noisy.data_ptr() + 4 * off, target.data_ptr() + 4 * off
    a.sample_stride, a.frame_stride = Ft * Cc * H * W, Cc * H * W
    a.sigmas, a.B, a.F, a.C, a.HW, a.weight = sigmas.data_ptr(), B, Fr, Cc, H * W, weight
    a.dpred, a.dpred_ld = dpred.data_ptr(), dpred.stride(0)
    a.workspace, a.loss, a.accumulate = ws.data_ptr(), loss.data_ptr(), int(accumulate)
    _lib.check(_lib.lib().pt_edm_loss(C.addressof(a), _sp()), "pt_edm_loss")


# ---------------------------------------------------------------------------------------------------------------
# block-level forward + backward
# ---------------------------------------------------------------------------------------------------------------
class FeedForwardTrainer:
    """out = x + Linear2(GEGLU(Linear1(LayerNorm(x)))) — BasicTransformerBlock.ff behind norm3 (modified_svd.py:100-107),
    forward on the inference kernels (unfused GEGLU so that the pre-activations exist), backward on the kernels above."""

    def __init__(self, ln_w, ln_b, w1, b1, w2, b2):
        self.ln_w, self.ln_b, self.w1, self.b1, self.w2, self.b2 = ln_w, ln_b, w1, b1, w2, b2
        self.saved: Dict[str, torch.Tensor] = {}

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        rows, Cc = x.shape
        dev = x.device
        ln = torch.empty_like(x)
        ops.LayerNorm(x, ln, self.ln_w, self.ln_b).launch(_sp())
        h = torch.empty(rows, self.w1.shape[0], device=dev, dtype=BF16)
        ops.Gemm(ln, self.w1, h, bias=self.b1).launch(_sp())
        act = geglu_forward(h)
        out = torch.empty_like(x)
        ops.Gemm(act, self.w2, out, bias=self.b2, res1=x).launch(_sp())
        self.saved = dict(x=x, ln=ln, h=h, act=act)
        return out

    def backward(self, dout: torch.Tensor):
        s = self.saved
        grads = {}
        grads["w2"] = wgrad(dout, s["act"])
        grads["b2"] = colsum(dout)[0]
        dact = linear_dgrad(dout, self.w2)
        dh = geglu_backward(s["h"], dact)
        grads["w1"] = wgrad(dh, s["ln"])
        grads["b1"] = colsum(dh)[0]
        dln = linear_dgrad(dh, self.w1)
        dx = dout.clone()                                    # the residual branch
        _, dgb = layernorm_backward(s["x"], dln, self.ln_w, accumulate_into=dx)
        grads["ln_w"], grads["ln_b"] = dgb[0], dgb[1]
        return dx, grads


class ResBlockTrainer:
    """SpatioTemporalResBlock (wiring as posetraj_b200/engine.py NetPlan.resblock) forward + backward.
    Parameters come as a dict of the diffusers names under the block prefix: spatial_res_block.{norm1,conv1,
    time_emb_proj,norm2,conv2,conv_shortcut}, temporal_res_block.{norm1,conv1,time_emb_proj,norm2,conv2},
    time_mixer.mix_factor.  `temb_s` / `temb_t` are the per-batch-row time-embedding projections [B, C] (fp32) the
    forward adds after conv1 (their gradients are returned as d_temb_s / d_temb_t)."""

    def __init__(self, params: Dict[str, torch.Tensor], *, B: int, F: int, H: int, W: int, eps: float):
        self.p, self.B, self.F, self.H, self.W, self.eps = params, B, F, H, W, eps
        self.n = B * F
        dev = next(iter(params.values())).device
        g = lambda k: params[k]
        f32 = lambda k: g(k).to(F32).contiguous()

        def conv3(k):
            w = g(k).to(F32)
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(BF16).contiguous()

        def tconv(k):
            w = g(k).to(F32)
            co, ci = w.shape[:2]
            return w.reshape(co, ci, 3).permute(0, 2, 1).reshape(co, 3 * ci).to(BF16).contiguous()

        s, t = "spatial_res_block.", "temporal_res_block."
        self.w = dict(c1=conv3(s + "conv1.weight"), c2=conv3(s + "conv2.weight"), t1=tconv(t + "conv1.weight"),
                      t2=tconv(t + "conv2.weight"))
        self.has_sc = (s + "conv_shortcut.weight") in params
        if self.has_sc:
            self.w["sc"] = g(s + "conv_shortcut.weight").to(BF16).reshape(g(s + "conv_shortcut.weight").shape[0], -1).contiguous()
        self.v = {k: f32(k) for k in params if k.endswith(("bias", "norm1.weight", "norm2.weight"))}
        self.alpha = float(torch.sigmoid(g("time_mixer.mix_factor").to(F32).reshape(-1)[0]).item())
        self.stats = torch.zeros((2 * self.n + 4 * ops.NUM_SMS + 64) * 64 + 1024, device=dev, dtype=torch.float64)
        self.saved: Dict[str, torch.Tensor] = {}

    def _gn(self, x, key, rows_per_stat, halo):
        n_img = x.shape[0] // (self.H * self.W)
        rows = n_img * (self.H + 1) * (self.W + 1) if halo else x.shape[0]
        out = torch.zeros(rows, x.shape[1], device=x.device, dtype=BF16)
        ops.GroupNorm(x, out, self.v[key + ".weight"], self.v[key + ".bias"], self.stats, rows_per_stat=rows_per_stat, eps=self.eps,
                      silu=True, halo=(self.H, self.W) if halo else None).launch(_sp())
        return out

    def forward(self, x: torch.Tensor, temb_s: torch.Tensor, temb_t: torch.Tensor) -> torch.Tensor:
        B, Fr, H, W, n = self.B, self.F, self.H, self.W, self.n
        HW = H * W
        rows = n * HW
        s, t = "spatial_res_block.", "temporal_res_block."
        taps = ops.conv3x3_taps(W)
        cout = self.w["c1"].shape[0]
        dev = x.device
        g1 = self._gn(x, s + "norm1", HW, True)
        h1 = torch.empty(rows, cout, device=dev, dtype=BF16)
        ops.Gemm(g1, self.w["c1"], h1, taps=taps, bias=self.v[s + "conv1.bias"], rowvec=temb_s, rowvec_mode=1, rv=(Fr * HW, 1, 1),
                 halo=(H, W)).launch(_sp())
        g2 = self._gn(h1, s + "norm2", HW, True)
        if self.has_sc:
            sc = torch.empty(rows, cout, device=dev, dtype=BF16)
            ops.Gemm(x, self.w["sc"], sc, bias=self.v[s + "conv_shortcut.bias"]).launch(_sp())
        else:
            sc = x
        xs = torch.empty(rows, cout, device=dev, dtype=BF16)
        ops.Gemm(g2, self.w["c2"], xs, taps=taps, bias=self.v[s + "conv2.bias"], res1=sc, halo=(H, W)).launch(_sp())
        t1 = self._gn(xs, t + "norm1", Fr * HW, False)
        t2 = torch.empty(rows, cout, device=dev, dtype=BF16)
        ops.Gemm(t1, self.w["t1"], t2, batches=B, taps=(-HW, 0, HW), bias=self.v[t + "conv1.bias"], rowvec=temb_t, rowvec_mode=1,
                 rv=(Fr * HW, 1, 1)).launch(_sp())
        t3 = self._gn(t2, t + "norm2", Fr * HW, False)
        y = torch.empty(rows, cout, device=dev, dtype=BF16)      # temporal conv2 output (kept: d mix_factor needs it)
        ops.Gemm(t3, self.w["t2"], y, batches=B, taps=(-HW, 0, HW), bias=self.v[t + "conv2.bias"]).launch(_sp())
        out = torch.empty(rows, cout, device=dev, dtype=BF16)
        ops.Axpy(xs, y, out, 1.0 - self.alpha).launch(_sp())   # blend(xs, xs + y) = xs + (1 - alpha) y
        self.saved = dict(x=x, g1=g1, h1=h1, g2=g2, xs=xs, t1=t1, t2=t2, t3=t3, y=y)
        return out

    def backward(self, dout: torch.Tensor):
        B, Fr, H, W, n = self.B, self.F, self.H, self.W, self.n
        HW = H * W
        sv, grads = self.saved, {}
        s, t = "spatial_res_block.", "temporal_res_block."
        taps = ops.conv3x3_taps(W)
        ttaps = (-HW, 0, HW)
        one_m_a = 1.0 - self.alpha
        # AlphaBlender: out = xs + (1 - alpha) y, alpha = sigmoid(mix_factor)
        grads["time_mixer.mix_factor"] = dot(dout, sv["y"], scale=-self.alpha * (1.0 - self.alpha))
        dy = torch.empty_like(dout)
        ops.Axpy(torch.zeros_like(dout), dout, dy, one_m_a).launch(_sp())
        # temporal conv2
        grads[t + "conv2.weight"] = wgrad(dy, sv["t3"], taps=ttaps, batches=B)
        grads[t + "conv2.bias"] = colsum(dy)[0]
        dt3 = linear_dgrad(dy, self.w["t2"], taps=ttaps, batches=B)
        dt2, _, dgb = groupnorm_backward(sv["t2"], dt3, self.v[t + "norm2.weight"], self.v[t + "norm2.bias"], rows_per_stat=Fr * HW,
                                         eps=self.eps)
        grads[t + "norm2.weight"], grads[t + "norm2.bias"] = dgb[0], dgb[1]
        # temporal conv1 (+ time embedding row vector)
        grads[t + "conv1.weight"] = wgrad(dt2, sv["t1"], taps=ttaps, batches=B)
        grads[t + "conv1.bias"] = colsum(dt2)[0]
        grads["d_temb_t"] = colsum(dt2, groups=B)
        dt1 = linear_dgrad(dt2, self.w["t1"], taps=ttaps, batches=B)
        dxs_t, _, dgb = groupnorm_backward(sv["xs"], dt1, self.v[t + "norm1.weight"], self.v[t + "norm1.bias"], rows_per_stat=Fr * HW,
                                           eps=self.eps)
        grads[t + "norm1.weight"], grads[t + "norm1.bias"] = dgb[0], dgb[1]
        dxs = torch.empty_like(dout)
        ops.Axpy(dout, dxs_t, dxs, 1.0).launch(_sp())          # residual path + temporal branch
        # spatial conv2: gradient in the zero-haloed row space of its input
        dxs_h = to_halo(dxs, n, H, W)
        grads[s + "conv2.weight"] = wgrad(dxs_h, sv["g2"], taps=taps)
        grads[s + "conv2.bias"] = colsum(dxs)[0]
        dg2 = linear_dgrad(dxs_h, self.w["c2"], taps=taps)
        dh1, _, dgb = groupnorm_backward(sv["h1"], dg2, self.v[s + "norm2.weight"], self.v[s + "norm2.bias"], rows_per_stat=HW,
                                         eps=self.eps, halo=(H, W))
        grads[s + "norm2.weight"], grads[s + "norm2.bias"] = dgb[0], dgb[1]
        # spatial conv1 (+ time embedding row vector)
        dh1_h = to_halo(dh1, n, H, W)
        grads[s + "conv1.weight"] = wgrad(dh1_h, sv["g1"], taps=taps)
        grads[s + "conv1.bias"] = colsum(dh1)[0]
        grads["d_temb_s"] = colsum(dh1, groups=B)
        dg1 = linear_dgrad(dh1_h, self.w["c1"], taps=taps)
        dx, _, dgb = groupnorm_backward(sv["x"], dg1, self.v[s + "norm1.weight"], self.v[s + "norm1.bias"], rows_per_stat=HW,
                                        eps=self.eps, halo=(H, W))
        grads[s + "norm1.weight"], grads[s + "norm1.bias"] = dgb[0], dgb[1]
        # shortcut
        if self.has_sc:
            grads[s + "conv_shortcut.weight"] = wgrad(dxs, sv["x"])
            grads[s + "conv_shortcut.bias"] = colsum(dxs)[0]
            dx = linear_dgrad(dxs, self.w["sc"], accumulate_into=dx)
        else:
            dx2 = torch.empty_like(dx)
            ops.Axpy(dx, dxs, dx2, 1.0).launch(_sp())
            dx = dx2
        return dx, grads


# ---------------------------------------------------------------------------------------------------------------
# data-parallel step: gradient buckets + AdamW
# ---------------------------------------------------------------------------------------------------------------
def bucket_plan(sizes: Sequence[int], bucket_elems: int) -> List[List[int]]:
    """Greedy packing of parameter indices (in REVERSE order: the backward pass finishes the last layers first) into
    buckets of at most `bucket_elems` elements (a larger parameter gets a bucket of its own).  Pure function: every rank
    derives the same plan."""
    buckets, cur, cur_n = [], [], 0
    for i in reversed(range(len(sizes))):
        if cur and cur_n + sizes[i] > bucket_elems:
            buckets.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += sizes[i]
    if cur:
        buckets.append(cur)
    return buckets


class GradientBuckets:
    """Flat fp32 gradient buckets averaged over the data-parallel group (what DDP does for the reference,
    train...cam_concat.py:1165,1470).  `ready(i)` marks parameter i's gradient as written; when a bucket is complete its
    all-reduce is launched at once on a side stream (NCCL) and runs under the rest of the backward pass; `finish()` waits
    for all of them.  Works on gloo with CPU tensors too (host-logic tests)."""

    def __init__(self, sizes: Sequence[int], device, group=None, bucket_mb: float = 100.0):
        import torch.distributed as dist
        self.sizes, self.group = list(sizes), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = bucket_plan(self.sizes, int(bucket_mb * (1 << 20) / 4))
        self.flat = [torch.zeros(sum(self.sizes[i] for i in b), device=device, dtype=F32) for b in self.buckets]
        self.where = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for i in b:
                self.where[i] = (bi, off)
                off += self.sizes[i]
        self.pending = [set(b) for b in self.buckets]
        self.handles = []
        self.cuda = torch.device(device).type == "cuda"
        self.side = torch.cuda.Stream(device=device) if self.cuda else None

    def view(self, i: int) -> torch.Tensor:
        bi, off = self.where[i]
        return self.flat[bi][off: off + self.sizes[i]]

    def ready(self, i: int) -> None:
        import torch.distributed as dist
        bi, _ = self.where[i]
        self.pending[bi].discard(i)
        if self.pending[bi] or self.world == 1:
            return
        if self.cuda:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                self.handles.append(dist.all_reduce(self.flat[bi], group=self.group, async_op=True))
        else:
            self.handles.append(dist.all_reduce(self.flat[bi], group=self.group, async_op=True))

    def finish(self) -> None:
        for h in self.handles:
            h.wait()
        if self.cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.side)
        self.handles = []
        self.pending = [set(b) for b in self.buckets]
        # the optimizer divides by the world size (grad_scale), so the sum is all that is needed here


class AdamW:
    """torch.optim.AdamW semantics (the reference's optimizer, train...cam_concat.py:1113-1119) as one fused kernel per
    gradient bucket: fp32 master weights, moments, and the bf16 working copy the forward kernels read."""

    def __init__(self, buckets: GradientBuckets, master: List[torch.Tensor], work: Optional[List[torch.Tensor]] = None, *, lr: float,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        self.b, self.lr, self.betas, self.eps, self.wd = buckets, lr, betas, eps, weight_decay
        self.step_count = 0
        dev = buckets.flat[0].device
        self.master = [torch.zeros_like(f) for f in buckets.flat]
        self.work = [torch.zeros(f.numel(), device=dev, dtype=BF16) for f in buckets.flat] if work is not None else None
        for i, p in enumerate(master):
            bi, off = buckets.where[i]
            self.master[bi][off: off + p.numel()].copy_(p.reshape(-1).to(F32))
        self.m = [torch.zeros_like(f) for f in buckets.flat]
        self.v = [torch.zeros_like(f) for f in buckets.flat]

    def param(self, i: int) -> torch.Tensor:
        bi, off = self.b.where[i]
        return self.master[bi][off: off + self.b.sizes[i]]

    def step(self) -> None:
        self.step_count += 1
        for bi, g in enumerate(self.b.flat):
            a = _lib.PtAdamWArgs()
            a.master, a.grad, a.m, a.v = self.master[bi].data_ptr(), g.data_ptr(), self.m[bi].data_ptr(), self.v[bi].data_ptr()
            if self.work is not None:
                a.work = self.work[bi].data_ptr()
            a.n = g.numel()
            a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.wd
            a.grad_scale = 1.0 / self.b.world
            a.step = self.step_count
            _lib.check(_lib.lib().pt_adamw(C.addressof(a), _sp()), "pt_adamw")


# ---------------------------------------------------------------------------------------------------------------
# whole-network backward: the remaining operators (csrc/train_attn.cu, csrc/train_misc.cu)
# ---------------------------------------------------------------------------------------------------------------
def attention_spatial_backward(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, *, n_img: int,
                               heads: int) -> torch.Tensor:
    """d(Q | K | V) of the per-image self-attention (oracle/backward.py attention_backward); `lse` is what the forward
    kernel wrote (ops.AttnSpatial(lse=...)), `out` its output."""
    _check(qkv, BF16), _check(out, BF16), _check(dout, BF16), _check(lse, F32)
    rows, c3 = qkv.shape
    Cc, S = c3 // 3, rows // n_img
    delta = torch.empty(n_img * heads * S, device=qkv.device, dtype=F32)
    _lib.check(_lib.lib().pt_attention_delta(out.data_ptr(), out.stride(0), dout.data_ptr(), dout.stride(0), delta.data_ptr(), rows, S,
                                             heads, _sp()), "pt_attention_delta")
    dqkv = torch.empty(rows, c3, device=qkv.device, dtype=BF16)
    a = _lib.PtAttnSpatialBwdArgs()
    a.qkv, a.ld, a.dout, a.dout_ld = qkv.data_ptr(), qkv.stride(0), dout.data_ptr(), dout.stride(0)
    a.lse, a.delta, a.dqkv, a.dld = lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), dqkv.stride(0)
    a.S, a.heads, a.C, a.n_img = S, heads, Cc, n_img
    if S >= 256 and qkv.stride(0) % 8 == 0 and dout.stride(0) % 8 == 0:
        # tcgen05 dK / dV kernel: TMA tiles of the operands (keep the descriptors alive until the call has been issued)
        tm_q = _lib.encode_tensormap(qkv.data_ptr(), [c3, S, n_img], [qkv.stride(0) * 2, qkv.stride(0) * 2 * S], [64, 128, 1])
        tm_d = _lib.encode_tensormap(dout.data_ptr(), [Cc, S, n_img], [dout.stride(0) * 2, dout.stride(0) * 2 * S], [64, 128, 1])
        a.tmap_qkv, a.tmap_dout = C.addressof(tm_q), C.addressof(tm_d)
    _lib.check(_lib.lib().pt_attention_spatial_bwd(C.addressof(a), _sp()), "pt_attention_spatial_bwd")
    return dqkv


def attention_temporal_backward(qkv: torch.Tensor, dout: torch.Tensor, *, batch: int, frames: int, hw: int, heads: int) -> torch.Tensor:
    _check(qkv, BF16), _check(dout, BF16)
    dqkv = torch.empty_like(qkv)
    a = _lib.PtAttnTemporalBwdArgs()
    a.qkv, a.ld, a.dout, a.dout_ld, a.dqkv, a.dld = qkv.data_ptr(), qkv.stride(0), dout.data_ptr(), dout.stride(0), dqkv.data_ptr(), dqkv.stride(0)
    a.B, a.F, a.HW, a.heads, a.C = batch, frames, hw, heads, qkv.shape[1] // 3
    _lib.check(_lib.lib().pt_attention_temporal_bwd(C.addressof(a), _sp()), "pt_attention_temporal_bwd")
    return dqkv


def upsample_backward(dout: torch.Tensor, *, n: int, H: int, W: int, halo: bool, scale: int) -> torch.Tensor:
    """Gradient of ops.Upsample2x: compact [n*H*W, C] from the (haloed) scale x scale output gradient."""
    _check(dout, BF16)
    dx = torch.empty(n * H * W, dout.shape[1], device=dout.device, dtype=BF16)
    a = _lib.PtUpsampleArgs()
    a.x, a.ld, a.out, a.out_ld = dx.data_ptr(), dx.stride(0), dout.data_ptr(), dout.stride(0)
    a.n, a.H, a.W, a.C, a.halo, a.scale = n, H, W, dout.shape[1], int(halo), scale
    _lib.check(_lib.lib().pt_upsample2x_bwd(C.addressof(a), _sp()), "pt_upsample2x_bwd")
    return dx


def dilate2x(src: torch.Tensor, *, n: int, H: int, W: int, src_halo: bool) -> torch.Tensor:
    """stride-2 conv output gradient -> zero-haloed [n*(H+1)*(W+1), C] rows of the conv's INPUT space."""
    _check(src, BF16)
    dst = torch.empty(n * (H + 1) * (W + 1), src.shape[1], device=src.device, dtype=BF16)
    _lib.check(_lib.lib().pt_dilate2x(src.data_ptr(), src.stride(0), int(src_halo), dst.data_ptr(), dst.stride(0), n, H, W, src.shape[1],
                                      _sp()), "pt_dilate2x")
    return dst


def zero_halo(x: torch.Tensor, *, n: int, H: int, W: int) -> torch.Tensor:
    _check(x, BF16)
    assert x.shape[0] == n * (H + 1) * (W + 1)
    _lib.check(_lib.lib().pt_zero_halo(x.data_ptr(), x.stride(0), n, H, W, x.shape[1], _sp()), "pt_zero_halo")
    return x


def silu_forward(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _check(x, BF16)
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.lib().pt_silu_fwd(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1], _sp()), "pt_silu_fwd")
    return out


def silu_backward(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    _check(x, BF16), _check(dy, BF16)
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().pt_silu_bwd(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), dx.data_ptr(), dx.stride(0), x.shape[0],
                                      x.shape[1], _sp()), "pt_silu_bwd")
    return dx


def small_linear_backward(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, *, act_in_silu: bool = False, want_dx: bool = True,
                          dw: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None, accumulate_w: bool = False,
                          dx: Optional[torch.Tensor] = None, accumulate_dx: bool = False):
    """Backward of ops.SmallLinear (no output activation): returns (dx | None, dw, db) in fp32."""
    _check(x, F32), _check(w, BF16), _check(dy, F32)
    M, K = x.shape
    N = w.shape[0]
    assert dy.shape == (M, N) and w.shape[1] == K
    if dw is None:
        dw = torch.empty(N, K, device=x.device, dtype=F32)
        db = torch.empty(N, device=x.device, dtype=F32)
        accumulate_w = False
    if want_dx and dx is None:
        dx = torch.empty(M, K, device=x.device, dtype=F32)
        accumulate_dx = False
    a = _lib.PtSmallLinearBwdArgs()
    a.x, a.x_ld, a.w, a.w_ld, a.dy, a.dy_ld = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), dy.data_ptr(), dy.stride(0)
    a.M, a.N, a.K, a.act_in_silu = M, N, K, int(act_in_silu)
    if want_dx:
        ws = torch.empty(_lib.lib().pt_small_linear_bwd_workspace_bytes(M, N, K), device=x.device, dtype=torch.uint8)
        a.dx, a.dx_ld, a.accumulate_dx, a.dx_workspace = dx.data_ptr(), dx.stride(0), int(accumulate_dx), ws.data_ptr()
    a.dw, a.db, a.accumulate_w = dw.data_ptr(), (db.data_ptr() if db is not None else None), int(accumulate_w)
    _lib.check(_lib.lib().pt_small_linear_bwd(C.addressof(a), _sp()), "pt_small_linear_bwd")
    return (dx if want_dx else None), dw, db


def colsum_grouped(x: torch.Tensor, *, groups: int, mode: int, ga: int, gb: int = 1, gc: int = 1, scale: float = 1.0,
                   out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """fp32 [groups, C] sums of bf16 rows by group (PtColsumGroupedArgs: mode 1 r/ga, 2 the rowvec_mode-2 map, 3 (r/ga)%gc)."""
    _check(x, BF16)
    rows, Cc = x.shape
    if out is None:
        out = torch.zeros(groups, Cc, device=x.device, dtype=F32)
        accumulate = False
    ws = torch.empty(_lib.lib().pt_colsum_grouped_workspace_bytes(rows, groups, Cc), device=x.device, dtype=torch.uint8)
    a = _lib.PtColsumGroupedArgs()
    a.x, a.ld, a.rows, a.C, a.groups, a.mode, a.ga, a.gb, a.gc = x.data_ptr(), x.stride(0), rows, Cc, groups, mode, ga, gb, gc
    a.scale, a.out, a.out_ld, a.accumulate, a.workspace = scale, out.data_ptr(), out.stride(0), int(accumulate), ws.data_ptr()
    _lib.check(_lib.lib().pt_colsum_grouped(C.addressof(a), _sp()), "pt_colsum_grouped")
    return out
