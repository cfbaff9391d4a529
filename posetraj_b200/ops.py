"""Launch descriptors for the sm_100a kernels (host side of the C ABI).

Each class pre-builds the C argument struct once (pointers into caller-owned torch buffers, TMA
descriptors) so that running it is a single ctypes call; the engine keeps lists of them per model
and replays them every denoise step.  Tensors are only used for their device pointers, shapes and
strides — the math is in csrc/*.cu.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import PT_DT_BF16, PT_DT_F32

NUM_SMS = 148


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(stream=None) -> int:
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


# ---------------------------------------------------------------------------------------------------------------
# Tile configuration: (cta_pair, block_n) per GEMM shape.
#   1. shapes of the configs[1] denoise step measured on B200 by tools/gemm_sweep.py (round 2, after the warp-uniform
#      issue loops: gpurun_out/r2f_sweep.log, copied to profiles/r2c_gemm_sweep.txt).  The CTA pair now wins every 3x3
#      conv and most K >= 640 layers; in round 1 the ~120-cycle issue waterfall per MMA had hidden that.
#   2. otherwise the argmin of a per-tile cost model fitted to that sweep by tools/fit_cost_model.py (11 % rms log
#      error; its pick is within 2.1 % of the measured optimum on average, 13 % at worst):
#      time = launch + max tiles per CTA(-pair) x (k_iters x k-step + epilogue).
# ---------------------------------------------------------------------------------------------------------------
TUNED = {
    # (rows, n_out, K per tap, taps, geglu): (cta_pair, block_n)
    (80640, 320, 320, 1, 0): (False, 160),    # 56.3 us
    (80640, 960, 320, 1, 0): (False, 192),    # 74.8 us
    (80640, 1280, 320, 1, 1): (False, 256),   # 146.4 us
    (80640, 320, 1280, 1, 0): (False, 160),   # 78.8 us
    (20160, 640, 640, 1, 0): (False, 128),    # 35.8 us
    (20160, 1920, 640, 1, 0): (True, 224),    # 54.3 us
    (20160, 2560, 640, 1, 1): (True, 256),    # 97.2 us
    (20160, 640, 2560, 1, 0): (True, 160),    # 64.5 us
    (5040, 1280, 1280, 1, 0): (False, 128),   # 27.6 us
    (5040, 3840, 1280, 1, 0): (True, 224),    # 41.9 us
    (5040, 5120, 1280, 1, 1): (True, 256),    # 85.0 us
    (5040, 1280, 5120, 1, 0): (True, 192),    # 58.4 us
    (1260, 1280, 1280, 1, 0): (False, 96),    # 15.4 us
    (1260, 5120, 1280, 1, 1): (True, 192),    # 33.8 us
    (1260, 1280, 5120, 1, 0): (True, 96),     # 27.6 us
    (83804, 320, 320, 9, 0): (True, 160),     # 101.4 us  (1 466 TFLOP/s)
    (21756, 640, 640, 9, 0): (True, 160),     # 103.4 us
    (5852, 1280, 1280, 9, 0): (True, 224),    # 113.7 us
    (1680, 1280, 1280, 9, 0): (True, 128),    # 46.1 us
    (83804, 320, 640, 9, 0): (True, 160),     # 193.5 us
    (21756, 640, 1280, 9, 0): (True, 160),    # 209.5 us
    (5852, 1280, 2560, 9, 0): (True, 224),    # 222.3 us
    (80640, 320, 320, 3, 0): (False, 160),    # 68.5 us
    (20160, 640, 640, 3, 0): (True, 224),     # 54.2 us
    (5040, 1280, 1280, 3, 0): (True, 192),    # 46.1 us
    (1260, 1280, 1280, 3, 0): (False, 128),   # 25.6 us
}

_C_MMA, _C_FLOOR, _C_B, _C_E0, _C_E1, _C_FLOOR_P, _C_B_P, _C_LAUNCH = 1.1392, 108.64, 0.785, 199.11, 9.5858, 139.68, 0.352, 12877.0


def tile_cost_ns(rows: int, batches: int, n_out: int, k_iters: int, geglu: bool, has_res: bool, pair: bool, bn: int) -> float:
    tm = 256 if pair else 128
    m_tiles = batches * math.ceil(rows / batches / tm)
    per = bn // 2 if geglu else bn
    n_tiles = math.ceil(n_out / per)
    units = NUM_SMS // 2 if pair else NUM_SMS
    t_max = math.ceil(m_tiles * n_tiles / units)
    kstep = max(bn * _C_MMA, (_C_FLOOR_P + _C_B_P * bn) if pair else (_C_FLOOR + _C_B * bn))
    epi = max(0.0, _C_E0 + _C_E1 * per * (2.0 if geglu else 1.0) * (1.3 if has_res else 1.0))
    return _C_LAUNCH + t_max * (k_iters * kstep + epi)


def pick_config(rows: int, batches: int, n_out: int, k_per_tap: int, taps: int, geglu: bool, has_res: bool,
                block_n: Optional[int] = None, cta_pair: Optional[bool] = None):
    """(cta_pair, block_n) for a GEMM shape; explicit arguments are honoured."""
    if block_n is None and cta_pair is None:
        hit = TUNED.get((rows, n_out, k_per_tap, taps, int(geglu)))
        if hit is not None:
            return hit
    k_iters = taps * k_per_tap // 64
    best = None
    for pair in ((False, True) if cta_pair is None else (bool(cta_pair),)):
        cands = [256, 192, 128, 64] if geglu else [256, 224, 192, 160, 128, 96, 64, 32]
        if block_n is not None:
            cands = [block_n]
        for bn in cands:
            if pair and bn < 64:
                continue
            per = bn // 2 if geglu else bn
            if block_n is None and per > max(64 if pair else 32, ((n_out + 31) // 32) * 32):
                continue
            c = tile_cost_ns(rows, batches, n_out, k_iters, geglu, has_res, pair, bn)
            if best is None or c < best[0] - 1e-9:
                best = (c, pair, bn)
    if best is None:
        return (bool(cta_pair), 64 if cta_pair else 32)
    return best[1], best[2]


class Gemm:
    """D = epilogue(sum_taps A[m + shift] @ W_tap^T) — see PtGemmArgs in include/posetraj_b200.h."""

    def __init__(self, a0: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *,
                 a1: Optional[torch.Tensor] = None, batches: int = 1, taps: Sequence[int] = (0,),
                 n_out: Optional[int] = None, block_n: Optional[int] = None, geglu: bool = False,
                 bias: Optional[torch.Tensor] = None,
                 rowvec: Optional[torch.Tensor] = None, rowvec_mode: int = 0, rv=(1, 1, 1),
                 acc_scale: float = 1.0,
                 res1: Optional[torch.Tensor] = None, res1_scale: float = 1.0,
                 res2: Optional[torch.Tensor] = None, res2_scale: float = 1.0,
                 out2: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, aux_scale: float = 0.0,
                 halo: Optional[tuple] = None, ostride: int = 1, out_halo: bool = False,
                 act_silu=False, name: str = "gemm", alg_k: Optional[int] = None,
                 cta_pair: Optional[bool] = None, scatter: Optional[dict] = None,
                 acc_scale_dev: Optional[torch.Tensor] = None):
        assert a0.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
        assert a0.dim() == 2 and w.dim() == 2 and a0.stride(1) == 1 and w.stride(1) == 1
        self.name = name
        rows_total, k0 = a0.shape
        assert rows_total % batches == 0
        rpb = rows_total // batches
        k1 = 0
        if a1 is not None:
            assert a1.dtype == torch.bfloat16 and a1.shape[0] == rows_total and a1.stride(1) == 1
            k1 = a1.shape[1]
        if k0 % 64 or k1 % 64:
            raise ValueError("Gemm: K must be a multiple of 64")
        ntaps = len(taps)
        if w.shape[1] != ntaps * (k0 + k1):
            raise ValueError(f"Gemm {name}: weight K {w.shape[1]} != taps*(K0+K1) {ntaps * (k0 + k1)}")
        if geglu:
            gate_off = w.shape[0] // 2
            n_out = gate_off if n_out is None else n_out
        else:
            gate_off = 0
            n_out = w.shape[0] if n_out is None else n_out
        cta_pair, block_n = pick_config(rows_total, batches, n_out, k0 + k1, ntaps, geglu,
                                        res1 is not None or res2 is not None or out2 is not None, block_n, cta_pair)
        self.cta_pair = bool(cta_pair)
        self.block_n = block_n
        b_box_rows = block_n // 2   # each CTA of a pair stages half of B; a single CTA issues two such loads

        self.tm_a0 = _lib.encode_tensormap(a0.data_ptr(), [k0, rpb, batches],
                                           [a0.stride(0) * 2, a0.stride(0) * 2 * rpb], [64, 128, 1])
        self.tm_a1 = None
        if a1 is not None:
            self.tm_a1 = _lib.encode_tensormap(a1.data_ptr(), [k1, rpb, batches],
                                               [a1.stride(0) * 2, a1.stride(0) * 2 * rpb], [64, 128, 1])
        self.tm_b = _lib.encode_tensormap(w.data_ptr(), [w.shape[1], w.shape[0]], [w.stride(0) * 2],
                                          [64, b_box_rows])
        a = _lib.PtGemmArgs()
        a.tmap_a0 = C.addressof(self.tm_a0)
        a.tmap_a1 = C.addressof(self.tm_a1) if self.tm_a1 is not None else None
        a.tmap_b = C.addressof(self.tm_b)
        a.rows_per_batch = rpb
        a.batches = batches
        a.n_out = n_out
        a.k0_chunks = k0 // 64
        a.k1_chunks = k1 // 64
        a.num_taps = ntaps
        for i, s in enumerate(taps):
            a.tap_shift[i] = int(s)
        a.block_n = block_n
        a.geglu = 1 if geglu else 0
        a.gate_row_offset = gate_off
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.is_contiguous()
        a.bias = _ptr(bias)
        if rowvec is not None:
            assert rowvec.dtype == torch.float32 and rowvec.stride(-1) == 1
            a.rowvec = _ptr(rowvec)
            a.rowvec_ld = rowvec.stride(0) if rowvec.dim() == 2 else rowvec.numel()
            a.rowvec_mode = rowvec_mode if rowvec_mode else 1
            rv = tuple(int(v) for v in rv)
            a.rv_a, a.rv_b, a.rv_c = rv[:3]
            if len(rv) == 5:
                a.rv_mod, a.rv_off = rv[3], rv[4]
        a.acc_scale = acc_scale
        if acc_scale_dev is not None:
            # device scalar multiplied into acc_scale at run time (conditioning_scale: survives CUDA-graph replay)
            assert acc_scale_dev.dtype == torch.float32 and acc_scale_dev.numel() == 1 and not geglu
            a.acc_scale_ptr = acc_scale_dev.data_ptr()
        out_rows = out.shape[0]
        if scatter is not None:
            # fused all-to-all: `out` is this rank's destination buffer in the OTHER sharding; the rows this GEMM
            # produces (and its residual operands) live in the source layout of B * J * S rows
            out_rows = scatter["B"] * scatter["J"] * scatter["S"]
        for r in (res1, res2):
            if r is not None:
                assert r.dtype == torch.bfloat16 and r.stride(1) == 1 and r.shape[0] == out_rows
        if res1 is not None and res2 is not None:
            assert res1.stride(0) == res2.stride(0)
        a.res1 = _ptr(res1)
        a.res2 = _ptr(res2)
        a.res1_scale = res1_scale
        a.res2_scale = res2_scale
        a.res_ld = res1.stride(0) if res1 is not None else (res2.stride(0) if res2 is not None else 0)
        assert out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
        a.out = out.data_ptr()
        a.out_ld = out.stride(0)
        a.out_dtype = PT_DT_BF16 if out.dtype == torch.bfloat16 else PT_DT_F32
        if out2 is not None:
            assert out2.dtype == out.dtype and out2.stride(0) == out.stride(0)
            a.out2 = out2.data_ptr()
            if aux is not None:       # aux None: out2 = val (a second copy of the output that outlives the pooled one)
                assert aux.dtype == torch.bfloat16 and aux.stride(0) == out.stride(0)
                a.aux = aux.data_ptr()
                a.aux_scale = aux_scale
        if halo is not None:
            h, w_ = halo  # unpadded image height/width of the A row space
            a.map_mode = 1
            a.pW1 = w_ + 1
            a.pH1 = h + 1
            a.ostride = ostride
            a.oH = (h + ostride - 1) // ostride
            a.oW = (w_ + ostride - 1) // ostride
            n_img = rows_total // ((h + 1) * (w_ + 1))
            assert n_img * (h + 1) * (w_ + 1) == rows_total
            a.out_halo = 1 if out_halo else 0
            if out_halo:
                assert out_rows == n_img * (a.oH + 1) * (a.oW + 1), (out_rows, n_img, a.oH, a.oW)
            else:
                assert out_rows == n_img * a.oH * a.oW, (out_rows, n_img, a.oH, a.oW)
        else:
            a.map_mode = 0
            assert out_rows == rows_total
        if scatter is not None:
            assert out2 is None and out.dtype == torch.bfloat16 and not out_halo
            starts, counts, peers = scatter["starts"], scatter["counts"], scatter["peers"]
            assert len(starts) == len(counts) == len(peers) <= 8 and scatter["mode"] in (1, 2)
            a.scatter_mode, a.sc_world = scatter["mode"], len(peers)
            a.sc_J, a.sc_S = scatter["J"], scatter["S"]
            a.sc_kept_off, a.sc_kept_total = scatter["kept_off"], scatter["kept_total"]
            for q in range(len(peers)):
                a.sc_start[q], a.sc_count[q], a.sc_peer[q] = starts[q], counts[q], peers[q]
        a.act_silu = int(act_silu)   # 0 none, 1 / True SiLU, 2 GELU (erf), 3 quick-GELU
        a.cta_pair = 1 if cta_pair else 0
        # algorithmic FLOPs (bench.py roofline): true output pixels x true (un-padded) K
        valid_rows = out_rows if not (halo is not None and out_halo) else n_img * a.oH * a.oW
        self.kind = "gemm"
        self.alg_flops = 2.0 * valid_rows * n_out * (2 if geglu else 1) * (alg_k if alg_k is not None else ntaps * (k0 + k1))
        # algorithmic bytes (read-once / write-once): the A rows once whatever the tap count, the weights, every epilogue
        # operand and output
        esz = out.element_size()
        self.alg_bytes = float(rows_total * (k0 + k1) * 2 + w.numel() * 2 + valid_rows * n_out * esz
                               + sum(valid_rows * n_out * 2 for t in (res1, res2, aux) if t is not None)
                               + (valid_rows * n_out * esz if out2 is not None else 0))
        self.args = a
        self._keep = (a0, a1, w, out, bias, rowvec, res1, res2, out2, aux, acc_scale_dev)
        self._argp = C.addressof(a)
        # semantic record of the launch (posetraj_b200.train_engine differentiates op lists through it)
        self.io = SimpleNamespace(a0=a0, a1=a1, w=w, out=out, bias=bias, rowvec=rowvec, rowvec_mode=(rowvec_mode or 1) if rowvec is not None else 0,
                                  rv=tuple(int(v) for v in rv), acc_scale=acc_scale, acc_scale_dev=acc_scale_dev, res1=res1,
                                  res1_scale=res1_scale, res2=res2, res2_scale=res2_scale, out2=out2, aux=aux, aux_scale=aux_scale,
                                  taps=tuple(int(t) for t in taps), batches=batches, halo=halo, ostride=ostride, out_halo=out_halo,
                                  act=int(act_silu), geglu=bool(geglu), n_out=n_out, scatter=scatter, mix=None)

    def launch(self, stream_ptr: int) -> None:
        _lib.check(_lib.lib().pt_gemm(self._argp, stream_ptr), self.name)


class FusedMlp:
    """out = acc_scale * (GEGLU(x W1^T + b1) W2^T + b2) + res1_scale*res1 + res2_scale*res2 in one kernel (PtMlpArgs)."""

    @staticmethod
    def supported(C: int) -> bool:
        return C % 64 == 0 and 64 <= C <= 320

    def __init__(self, x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                 out: torch.Tensor, *, acc_scale: float = 1.0, res1: Optional[torch.Tensor] = None, res1_scale: float = 1.0,
                 res2: Optional[torch.Tensor] = None, res2_scale: float = 1.0, name: str = "mlp",
                 trace: Optional[torch.Tensor] = None):
        rows, Cc = x.shape
        hidden = w2.shape[1]
        assert x.dtype == w1.dtype == w2.dtype == out.dtype == torch.bfloat16
        assert x.stride(1) == 1 and w1.is_contiguous() and w2.is_contiguous() and out.stride(1) == 1
        assert w1.shape == (2 * hidden, Cc) and w2.shape == (Cc, hidden) and hidden == 4 * Cc and self.supported(Cc)
        assert b1.dtype == torch.float32 and b1.numel() == 2 * hidden and b2.dtype == torch.float32 and b2.numel() == Cc
        assert out.shape == (rows, Cc)
        self.name = name
        self.tm_x = _lib.encode_tensormap(x.data_ptr(), [Cc, rows], [x.stride(0) * 2], [64, 128])
        self.tm_w1 = _lib.encode_tensormap(w1.data_ptr(), [Cc, 2 * hidden], [w1.stride(0) * 2], [64, 64])
        self.tm_w2 = _lib.encode_tensormap(w2.data_ptr(), [hidden, Cc], [w2.stride(0) * 2], [64, Cc // 4])
        a = _lib.PtMlpArgs()
        a.tmap_x, a.tmap_w1, a.tmap_w2 = C.addressof(self.tm_x), C.addressof(self.tm_w1), C.addressof(self.tm_w2)
        a.rows, a.C, a.hidden = rows, Cc, hidden
        a.bias1, a.bias2 = b1.data_ptr(), b2.data_ptr()
        a.acc_scale = acc_scale
        for r in (res1, res2):
            if r is not None:
                assert r.dtype == torch.bfloat16 and r.stride(1) == 1 and r.shape == (rows, Cc)
        if res1 is not None and res2 is not None:
            assert res1.stride(0) == res2.stride(0)
        a.res1, a.res2 = _ptr(res1), _ptr(res2)
        a.res1_scale, a.res2_scale = res1_scale, res2_scale
        a.res_ld = res1.stride(0) if res1 is not None else (res2.stride(0) if res2 is not None else 0)
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        if trace is not None:
            assert trace.dtype == torch.int64 and trace.numel() >= 2 * 8 * 64
            a.trace = trace.data_ptr()
        self.kind = "mlp_geglu"
        self.alg_flops = 2.0 * rows * Cc * (2 * hidden) + 2.0 * rows * hidden * Cc
        self.alg_bytes = float(rows * Cc * 2 * (2 + sum(1 for t in (res1, res2) if t is not None)) + (w1.numel() + w2.numel()) * 2)
        self.args = a
        self._keep = (x, w1, b1, w2, b2, out, res1, res2)
        self._argp = C.addressof(a)

    def launch(self, stream_ptr: int) -> None:
        _lib.check(_lib.lib().pt_mlp_geglu(self._argp, stream_ptr), self.name)


def conv3x3_taps(w_img: int) -> list[int]:
    """Row shifts of the 9 taps (ky, kx row-major) in the zero-haloed image space of width w_img + 1."""
    return [dy * (w_img + 1) + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


class CfgEuler:
    def __init__(self, *, noise_pred: Optional[torch.Tensor], latents: torch.Tensor, guidance: torch.Tensor,
                 sigmas: torch.Tensor, step_index: torch.Tensor, next_in: Optional[torch.Tensor] = None,
                 image_latents: Optional[torch.Tensor] = None, next_padded: bool = True, mode: int = 0,
                 pred_nchw_f32: bool = False, single_pred: bool = False, row_begin: int = 0, row_count: int = 2):
        F_, Cc, H, W = latents.shape[-4:]
        assert latents.dtype == torch.float32 and latents.is_contiguous()
        assert guidance.dtype == torch.float32 and sigmas.dtype == torch.float32 and step_index.dtype == torch.int32
        a = _lib.PtCfgEulerArgs()
        a.noise_pred = _ptr(noise_pred)
        if noise_pred is not None and not pred_nchw_f32:
            assert noise_pred.dtype == torch.bfloat16 and noise_pred.dim() == 2
            a.pred_ld = noise_pred.stride(0)
        if pred_nchw_f32:
            assert noise_pred.dtype == torch.float32 and noise_pred.is_contiguous()
        a.pred_nchw_f32 = 1 if pred_nchw_f32 else 0
        a.latents = latents.data_ptr()
        a.guidance = guidance.data_ptr()
        a.sigmas = sigmas.data_ptr()
        a.step_index = step_index.data_ptr()
        a.F, a.C, a.H, a.W = F_, Cc, H, W
        if next_in is not None:
            assert next_in.dtype == torch.bfloat16 and next_in.dim() == 2
            assert image_latents is not None and image_latents.dtype == torch.float32 and image_latents.is_contiguous()
            a.next_in = next_in.data_ptr()
            a.image_latents = image_latents.data_ptr()
            a.next_ld = next_in.stride(0)
            a.next_padded = 1 if next_padded else 0
        a.mode = mode
        a.single_pred = 1 if single_pred else 0
        a.row_begin, a.row_count = row_begin, row_count
        self.kind = "cfg_euler"
        self.name = "pt_cfg_euler_step"
        self.alg_flops = 0.0
        # F*C*H*W*(2*2 + 4 + 4) B, plus the fused bf16 write of the next model input (2 rows x 2C ch)
        self.alg_bytes = F_ * Cc * H * W * (2 * 2 + 4 + 4.0) + (2 * F_ * H * W * 2 * Cc * 2.0 if next_in is not None else 0.0)
        self.args = a
        self._keep = (noise_pred, latents, guidance, sigmas, step_index, next_in, image_latents)
        self._argp = C.addressof(a)

    def launch(self, stream_ptr: int) -> None:
        _lib.check(_lib.lib().pt_cfg_euler_step(self._argp, stream_ptr), "pt_cfg_euler_step")


class _Op:
    """Base: holds a ctypes args struct + the C entry point; launch() is one ctypes call."""
    fn_name = ""
    kind = "misc"
    alg_flops = 0.0   # algorithmic FLOPs / bytes of one launch, for bench.py's roofline
    alg_bytes = 0.0

    def _finish(self, args, keep, name=None):
        self.args = args
        self._keep = keep
        self._argp = C.addressof(args)
        self._fn = getattr(_lib.lib(), self.fn_name)
        self.name = name or self.fn_name

    def launch(self, stream_ptr: int) -> None:
        _lib.check(self._fn(self._argp, stream_ptr), self.name)


class GroupNorm(_Op):
    fn_name = "pt_groupnorm"

    def __init__(self, x0, out, gamma, beta, stats, *, rows_per_stat, eps, silu=True, x1=None,
                 halo: Optional[tuple] = None, name=None, mode: int = 0, sums: Optional[torch.Tensor] = None,
                 count: float = 0.0, sums_peers: Optional[Sequence[int]] = None):
        a = _lib.PtGroupNormArgs()
        a.mode = mode
        if sums_peers is not None:
            assert mode == 2 and 1 <= len(sums_peers) <= 8
            a.n_peers = len(sums_peers)
            for q, ptr in enumerate(sums_peers):
                a.sums_peers[q] = ptr
        if mode != 0:
            assert sums is not None and sums.dtype == torch.float64 and sums.is_contiguous()
            a.sums, a.count = sums.data_ptr(), float(count)
        assert x0.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and stats.dtype == torch.float64
        assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
        rows = x0.shape[0]
        assert rows % rows_per_stat == 0
        a.x0 = x0.data_ptr()
        a.c0, a.ld0 = x0.shape[1], x0.stride(0)
        if x1 is not None:
            assert x1.dtype == torch.bfloat16 and x1.shape[0] == rows
            a.x1 = x1.data_ptr()
            a.c1, a.ld1 = x1.shape[1], x1.stride(0)
        assert gamma.numel() == a.c0 + a.c1
        a.rows_per_stat = rows_per_stat
        a.num_stat = rows // rows_per_stat
        need = _lib.lib().pt_groupnorm_workspace_bytes(a.num_stat, rows_per_stat, a.c0 + a.c1)
        assert need > 0 and stats.numel() * 8 >= need, (need, stats.numel())
        a.stats = stats.data_ptr()
        a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
        a.eps = eps
        a.silu = 1 if silu else 0
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        if halo is not None:
            a.halo, a.H, a.W = 1, halo[0], halo[1]
            n_img = rows // (halo[0] * halo[1])
            assert out.shape[0] == n_img * (halo[0] + 1) * (halo[1] + 1)
        else:
            assert out.shape[0] == rows
        self.kind = "groupnorm"
        self.alg_bytes = 2.0 * rows * (a.c0 + a.c1) * 2 * (0.5 if mode else 1.0)
        self.io = SimpleNamespace(x0=x0, x1=x1, out=out, gamma=gamma, beta=beta, rows_per_stat=rows_per_stat, eps=eps, silu=bool(silu),
                                  halo=halo, mode=mode)
        self._finish(a, (x0, x1, out, gamma, beta, stats, sums), name)


class LayerNorm(_Op):
    fn_name = "pt_layernorm"

    def __init__(self, x, out, gamma, beta, *, eps=1e-5, addvec=None, hw=1, frames=1, sum_out=None, name=None):
        a = _lib.PtLayerNormArgs()
        assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
        a.x, a.ld = x.data_ptr(), x.stride(0)
        a.gamma, a.beta, a.eps = gamma.data_ptr(), beta.data_ptr(), eps
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.rows, a.C = x.shape
        if addvec is not None:
            assert addvec.dtype == torch.float32 and addvec.is_contiguous() and addvec.shape == (frames, x.shape[1])
            a.addvec, a.hw, a.F = addvec.data_ptr(), hw, frames
            if sum_out is not None:
                assert sum_out.stride(0) == out.stride(0) and sum_out.dtype == torch.bfloat16
                a.sum_out = sum_out.data_ptr()
        self.kind = "layernorm"
        self.alg_bytes = (3.0 if sum_out is not None else 2.0) * x.shape[0] * x.shape[1] * 2
        self.io = SimpleNamespace(x=x, out=out, gamma=gamma, beta=beta, eps=eps, addvec=addvec, hw=hw, frames=frames, sum_out=sum_out)
        self._finish(a, (x, out, gamma, beta, addvec, sum_out), name)


class AttnSpatial(_Op):
    fn_name = "pt_attention_spatial"

    def __init__(self, qkv, out, *, n_img, heads, name=None, lse: Optional[torch.Tensor] = None):
        rows, c3 = qkv.shape
        Cc = c3 // 3
        S = rows // n_img
        assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and S * n_img == rows
        self.tm = _lib.encode_tensormap(qkv.data_ptr(), [c3, S, n_img], [qkv.stride(0) * 2, qkv.stride(0) * 2 * S],
                                        [64, 128, 1])
        a = _lib.PtAttnSpatialArgs()
        a.tmap_qkv = C.addressof(self.tm)
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.S, a.heads, a.C, a.n_img = S, heads, Cc, n_img
        if lse is not None:   # training: keep the per-row log-sum-exp for pt_attention_spatial_bwd
            assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == n_img * heads * S
            a.lse = lse.data_ptr()
        self.kind = "attn_spatial"
        self.alg_flops = 4.0 * S * S * 64 * heads * n_img
        self.alg_bytes = 4.0 * rows * Cc * 2
        self.io = SimpleNamespace(qkv=qkv, out=out, n_img=n_img, heads=heads, lse=lse)
        self._finish(a, (qkv, out, lse), name)


class AttnTemporal(_Op):
    fn_name = "pt_attention_temporal"

    def __init__(self, qkv, out, *, batch, frames, hw, heads, name=None):
        a = _lib.PtAttnTemporalArgs()
        assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
        assert qkv.shape[0] == batch * frames * hw
        a.qkv, a.ld = qkv.data_ptr(), qkv.stride(0)
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.B, a.F, a.HW, a.heads, a.C = batch, frames, hw, heads, qkv.shape[1] // 3
        self.kind = "attn_temporal"
        self.alg_flops = 4.0 * frames * frames * 64 * heads * batch * hw
        self.alg_bytes = 4.0 * qkv.shape[0] * a.C * 2
        self.io = SimpleNamespace(qkv=qkv, out=out, batch=batch, frames=frames, hw=hw, heads=heads)
        self._finish(a, (qkv, out), name)


class SmallLinear(_Op):
    fn_name = "pt_small_linear"

    def __init__(self, x, w, out, bias=None, *, act_in_silu=False, act_out_silu=False, accumulate=False, name=None):
        a = _lib.PtSmallLinearArgs()
        assert x.dtype == torch.float32 and out.dtype == torch.float32 and w.dtype == torch.bfloat16
        assert x.dim() == 2 and out.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1
        a.in_, a.in_ld = x.data_ptr(), x.stride(0)
        a.w, a.w_ld = w.data_ptr(), w.stride(0)
        a.bias = _ptr(bias)
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.M, a.K = x.shape
        a.N = w.shape[0]
        assert w.shape[1] == a.K and out.shape == (a.M, a.N)
        a.act_in_silu, a.act_out_silu, a.accumulate = int(act_in_silu), int(act_out_silu), int(accumulate)
        self.io = SimpleNamespace(x=x, w=w, out=out, bias=bias, act_in_silu=bool(act_in_silu), act_out_silu=bool(act_out_silu),
                                  accumulate=bool(accumulate))
        self._finish(a, (x, w, out, bias), name)


class SinCos(_Op):
    fn_name = "pt_timestep_sincos"

    def __init__(self, out, *, t=None, sigmas=None, step_index=None, name=None):
        a = _lib.PtSinCosArgs()
        assert out.dtype == torch.float32 and out.dim() == 2
        a.t, a.sigmas, a.step_index = _ptr(t), _ptr(sigmas), _ptr(step_index)
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.M, a.dim = out.shape
        if t is not None:
            assert t.dtype == torch.float32 and t.numel() == a.M
        self._finish(a, (out, t, sigmas, step_index), name)


class Upsample2x(_Op):
    fn_name = "pt_upsample2x"

    def __init__(self, x, out, *, n, H, W, halo=True, scale=2, name=None):
        a = _lib.PtUpsampleArgs()
        assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and x.shape[0] == n * H * W
        a.x, a.ld = x.data_ptr(), x.stride(0)
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
        a.n, a.H, a.W, a.C, a.halo, a.scale = n, H, W, x.shape[1], int(halo), scale
        exp_rows = n * (scale * H + 1) * (scale * W + 1) if halo else n * scale * scale * H * W
        assert out.shape[0] == exp_rows
        self.kind = "upsample"
        self.alg_bytes = (x.shape[0] + out.shape[0]) * x.shape[1] * 2.0
        self.io = SimpleNamespace(x=x, out=out, n=n, H=H, W=W, halo=bool(halo), scale=scale)
        self._finish(a, (x, out), name)


class ConvDirect(_Op):
    fn_name = "pt_conv3x3_direct"

    def __init__(self, x, w, bias, out, *, n, H, W, cin, cout, stride=1, silu=True, in_nchw_f32=False,
                 out_halo=False, name=None):
        a = _lib.PtConvDirectArgs()
        assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == 9 * cin * cout
        assert bias.dtype == torch.float32 and out.dtype == torch.bfloat16
        a.x = x.data_ptr()
        a.in_nchw_f32 = int(in_nchw_f32)
        if in_nchw_f32:
            assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() == n * cin * H * W
        else:
            assert x.dtype == torch.bfloat16 and x.shape[0] == n * H * W
            a.in_ld = x.stride(0)
        a.w, a.bias = w.data_ptr(), bias.data_ptr()
        a.out, a.out_ld, a.out_halo = out.data_ptr(), out.stride(0), int(out_halo)
        a.n, a.H, a.W, a.Cin, a.Cout, a.stride, a.silu = n, H, W, cin, cout, stride, int(silu)
        self._finish(a, (x, w, bias, out), name)


class Layout(_Op):
    """NCHW (fp32|bf16) <-> token-major bf16."""

    def __init__(self, nchw, tokens, *, to_tokens: bool, halo=False, name=None):
        self.fn_name = "pt_nchw_to_tokens" if to_tokens else "pt_tokens_to_nchw"
        a = _lib.PtLayoutArgs()
        assert nchw.is_contiguous() and nchw.dim() == 4 and nchw.dtype in (torch.float32, torch.bfloat16)
        assert tokens.dtype == torch.bfloat16 and tokens.dim() == 2
        a.nchw, a.tokens = nchw.data_ptr(), tokens.data_ptr()
        a.n, a.C, a.H, a.W = nchw.shape
        a.ld, a.halo, a.nchw_f32 = tokens.stride(0), int(halo), int(nchw.dtype == torch.float32)
        rows = a.n * ((a.H + 1) * (a.W + 1) if halo else a.H * a.W)
        assert tokens.shape[0] == rows and tokens.shape[1] >= a.C
        self._finish(a, (nchw, tokens), name)


class RowBlockCopy(_Op):
    """Copies row blocks between two token layouts: block i = `rows[i]` rows from src row `src_row[i]` to dst row `dst_row[i]`."""
    fn_name = "pt_row_block_copy"
    kind = "layout"

    def __init__(self, src, dst, src_rows, dst_rows, rows, name=None):
        a = _lib.PtRowBlockCopyArgs()
        assert src.dtype == torch.bfloat16 and dst.dtype == torch.bfloat16 and src.shape[1] == dst.shape[1]
        dev = src.device
        self.t_src = torch.tensor(src_rows, dtype=torch.int32, device=dev)
        self.t_dst = torch.tensor(dst_rows, dtype=torch.int32, device=dev)
        self.t_rows = torch.tensor(rows, dtype=torch.int32, device=dev)
        assert len(src_rows) == len(dst_rows) == len(rows) > 0
        a.src, a.dst = src.data_ptr(), dst.data_ptr()
        a.src_ld, a.dst_ld, a.cols = src.stride(0), dst.stride(0), src.shape[1]
        a.src_row, a.dst_row, a.rows = self.t_src.data_ptr(), self.t_dst.data_ptr(), self.t_rows.data_ptr()
        a.n_blocks = len(rows)
        self.alg_bytes = 2.0 * sum(rows) * src.shape[1] * 2
        self._finish(a, (src, dst), name)


class Axpy(_Op):
    """out = x + scale * y (bf16)."""
    fn_name = "pt_axpy_bf16"
    kind = "layout"

    def __init__(self, x, y, out, scale: float, name=None):
        a = _lib.PtAxpyArgs()
        assert x.dtype == y.dtype == out.dtype == torch.bfloat16 and x.shape == y.shape == out.shape
        a.x, a.y, a.out = x.data_ptr(), y.data_ptr(), out.data_ptr()
        a.ld_x, a.ld_y, a.ld_out = x.stride(0), y.stride(0), out.stride(0)
        a.rows, a.cols = x.shape
        a.scale = float(scale)
        self.alg_bytes = 3.0 * x.numel() * 2
        self.io = SimpleNamespace(x=x, y=y, out=out, scale=float(scale))
        self._finish(a, (x, y, out), name)


class GegluFwd:
    """out = value * gelu(gate) of h = (value | gate) as its own pass (training keeps the pre-activations)."""
    kind, alg_flops = "train_misc", 0.0

    def __init__(self, h, out, name="geglu"):
        assert h.dtype == out.dtype == torch.bfloat16 and h.shape[1] == 2 * out.shape[1] and h.shape[0] == out.shape[0]
        self.io = SimpleNamespace(h=h, out=out)
        self.name = name
        self.alg_bytes = 3.0 * out.numel() * 2

    def launch(self, stream_ptr: int) -> None:
        h, out = self.io.h, self.io.out
        _lib.check(_lib.lib().pt_geglu_fwd(h.data_ptr(), h.stride(0), out.data_ptr(), out.stride(0), h.shape[0], out.shape[1], stream_ptr),
                   self.name)


class SiluFwd:
    """out = silu(x) (bf16 rows) as its own pass (conditioning embedding in training)."""
    kind, alg_flops = "train_misc", 0.0

    def __init__(self, x, out, name="silu"):
        assert x.dtype == out.dtype == torch.bfloat16 and x.shape == out.shape
        self.io = SimpleNamespace(x=x, out=out)
        self.name = name
        self.alg_bytes = 2.0 * out.numel() * 2

    def launch(self, stream_ptr: int) -> None:
        x, out = self.io.x, self.io.out
        _lib.check(_lib.lib().pt_silu_fwd(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1], stream_ptr),
                   self.name)


class SoftmaxRows(_Op):
    """bf16 out[r, :cols] = softmax(fp32 x[r, :cols]) (VAE mid-block attention)."""
    fn_name = "pt_softmax_rows"
    kind = "vae_misc"

    def __init__(self, x, out, cols: int, name=None):
        a = _lib.PtSoftmaxArgs()
        assert x.dtype == torch.float32 and out.dtype == torch.bfloat16 and x.shape[0] == out.shape[0]
        assert x.stride(1) == 1 and out.stride(1) == 1 and cols <= x.shape[1] and cols <= out.shape[1]
        a.in_, a.out = x.data_ptr(), out.data_ptr()
        a.rows, a.cols, a.ld_in, a.ld_out = x.shape[0], cols, x.stride(0), out.stride(0)
        self.alg_bytes = x.shape[0] * cols * 6.0
        self._finish(a, (x, out), name)


class TimeConv3(_Op):
    """TemporalDecoder.time_conv_out on token-major fp32 rows -> NCHW fp32 frames."""
    fn_name = "pt_time_conv3"
    kind = "vae_misc"

    def __init__(self, x, w, bias, out, *, batch: int, frames: int, hw: int, name=None):
        a = _lib.PtTimeConvArgs()
        Cc = out.shape[1]
        assert x.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous() and x.stride(1) == 1
        assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == Cc * Cc * 3 and bias.numel() == Cc
        assert x.shape[0] == batch * frames * hw and out.shape[0] == batch * frames and out[0, 0].numel() == hw
        a.in_, a.ld, a.w, a.bias, a.out = x.data_ptr(), x.stride(0), w.data_ptr(), bias.data_ptr(), out.data_ptr()
        a.B, a.F, a.HW, a.C = batch, frames, hw, Cc
        self.alg_bytes = x.shape[0] * Cc * 8.0
        self._finish(a, (x, w, bias, out), name)


class BlurReflect(_Op):
    """One pass of the reference's separable Gaussian pre-filter (reflect padding) on fp32 planes."""
    fn_name = "pt_blur_reflect"
    kind = "clip_misc"

    def __init__(self, x, out, w, *, axis: int, name=None):
        a = _lib.PtBlurArgs()
        assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
        assert x.shape == out.shape and x.dim() == 3 and w.dtype == torch.float32 and w.is_contiguous()
        a.in_, a.out, a.w = x.data_ptr(), out.data_ptr(), w.data_ptr()
        a.planes, a.H, a.W = x.shape
        a.k, a.axis = w.numel(), axis
        self.alg_bytes = x.numel() * 8.0
        self._finish(a, (x, out, w), name)


class BicubicResize(_Op):
    """bicubic, align_corners=True, to S x S: fp32 planes and/or bf16 patch-embedding rows."""
    fn_name = "pt_bicubic_resize"
    kind = "clip_misc"

    def __init__(self, x, *, size: int, out_f32=None, out_patches=None, patch: int = 1, name=None):
        a = _lib.PtBicubicArgs()
        assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
        a.in_ = x.data_ptr()
        a.C, a.H, a.W = x.shape
        a.S, a.P = size, patch
        if out_f32 is not None:
            assert out_f32.dtype == torch.float32 and out_f32.is_contiguous() and out_f32.numel() == a.C * size * size
            a.out_f32 = out_f32.data_ptr()
        if out_patches is not None:
            assert out_patches.dtype == torch.bfloat16 and out_patches.stride(1) == 1
            assert out_patches.shape[0] == (size // patch) ** 2
            a.out_patches, a.ld = out_patches.data_ptr(), out_patches.stride(0)
        self._finish(a, (x, out_f32, out_patches), name)


class AttnSmall(_Op):
    """Self-attention for short sequences / any head_dim <= 128 (CLIP vision tower)."""
    fn_name = "pt_attention_small"
    kind = "clip_attn"

    def __init__(self, qkv, out, *, heads: int, name=None):
        a = _lib.PtAttnSmallArgs()
        S, c3 = qkv.shape
        assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and out.shape == (S, c3 // 3)
        assert qkv.stride(1) == 1 and out.stride(1) == 1 and (c3 // 3) % heads == 0
        a.qkv, a.ld, a.out, a.out_ld = qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0)
        a.S, a.heads, a.head_dim = S, heads, c3 // 3 // heads
        self.alg_flops = 4.0 * S * S * (c3 // 3)
        self._finish(a, (qkv, out), name)


class TorchOp:
    """A tiny torch-side op inside an op list that is replayed OUTSIDE the per-step graph (embedding staging)."""
    kind, alg_flops, alg_bytes = "misc", 0.0, 0.0

    def __init__(self, fn, name="torch_op"):
        self.fn, self.name = fn, name

    def launch(self, stream_ptr: int) -> None:
        self.fn()


class StepAdvance:
    def __init__(self, step_index: torch.Tensor):
        assert step_index.dtype == torch.int32
        self.step_index = step_index
        self.name = "pt_step_advance"
        self.kind, self.alg_flops, self.alg_bytes = "misc", 0.0, 0.0

    def launch(self, stream_ptr: int) -> None:
        _lib.check(_lib.lib().pt_step_advance(self.step_index.data_ptr(), stream_ptr), self.name)
