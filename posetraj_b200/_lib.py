"""ctypes binding of libposetraj_b200.so (include/posetraj_b200.h).

There is exactly one compute backend: the hand-written sm_100a library.  If it is missing or a call
fails, this module raises — there is no eager/CPU fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libposetraj_b200.so"


class PtTensorMap(C.Structure):
    _fields_ = [("opaque", C.c_uint64 * 16)]


class PtCfgEulerArgs(C.Structure):
    _fields_ = [
        ("noise_pred", C.c_void_p),
        ("pred_ld", C.c_int32),
        ("pred_nchw_f32", C.c_int32),
        ("latents", C.c_void_p),
        ("guidance", C.c_void_p),
        ("sigmas", C.c_void_p),
        ("step_index", C.c_void_p),
        ("F", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("next_in", C.c_void_p),
        ("image_latents", C.c_void_p),
        ("next_ld", C.c_int32),
        ("next_padded", C.c_int32),
        ("mode", C.c_int32),
        ("single_pred", C.c_int32),
        ("row_begin", C.c_int32),
        ("row_count", C.c_int32),
    ]


class PtGemmArgs(C.Structure):
    _fields_ = [
        ("tmap_a0", C.c_void_p),
        ("tmap_a1", C.c_void_p),
        ("tmap_b", C.c_void_p),
        ("rows_per_batch", C.c_int32),
        ("batches", C.c_int32),
        ("n_out", C.c_int32),
        ("k0_chunks", C.c_int32),
        ("k1_chunks", C.c_int32),
        ("num_taps", C.c_int32),
        ("tap_shift", C.c_int32 * 9),
        ("block_n", C.c_int32),
        ("geglu", C.c_int32),
        ("gate_row_offset", C.c_int32),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p),
        ("rowvec_ld", C.c_int32),
        ("rowvec_mode", C.c_int32),
        ("rv_a", C.c_int32), ("rv_b", C.c_int32), ("rv_c", C.c_int32),
        ("acc_scale", C.c_float),
        ("res1", C.c_void_p),
        ("res2", C.c_void_p),
        ("res1_scale", C.c_float), ("res2_scale", C.c_float),
        ("res_ld", C.c_int32),
        ("out", C.c_void_p),
        ("out_ld", C.c_int32),
        ("out_dtype", C.c_int32),
        ("out2", C.c_void_p),
        ("aux", C.c_void_p),
        ("aux_scale", C.c_float),
        ("map_mode", C.c_int32),
        ("pW1", C.c_int32), ("pH1", C.c_int32), ("ostride", C.c_int32), ("oW", C.c_int32), ("oH", C.c_int32),
        ("out_halo", C.c_int32), ("act_silu", C.c_int32), ("cta_pair", C.c_int32),
        ("rv_mod", C.c_int32), ("rv_off", C.c_int32),
        ("scatter_mode", C.c_int32), ("sc_world", C.c_int32), ("sc_J", C.c_int32), ("sc_S", C.c_int32),
        ("sc_kept_off", C.c_int32), ("sc_kept_total", C.c_int32),
        ("sc_start", C.c_int32 * 8), ("sc_count", C.c_int32 * 8), ("sc_peer", C.c_void_p * 8),
        ("acc_scale_ptr", C.c_void_p),
    ]



def _st(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


i32, f32, vp = C.c_int32, C.c_float, C.c_void_p

PtMlpArgs = _st("PtMlpArgs", [
    ("tmap_x", vp), ("tmap_w1", vp), ("tmap_w2", vp), ("rows", i32), ("C", i32), ("hidden", i32),
    ("bias1", vp), ("bias2", vp), ("acc_scale", f32), ("res1", vp), ("res2", vp),
    ("res1_scale", f32), ("res2_scale", f32), ("res_ld", i32), ("out", vp), ("out_ld", i32), ("trace", vp)])

PtGroupNormArgs = _st("PtGroupNormArgs", [
    ("x0", vp), ("x1", vp), ("c0", i32), ("c1", i32), ("ld0", i32), ("ld1", i32),
    ("rows_per_stat", i32), ("num_stat", i32), ("stats", vp), ("gamma", vp), ("beta", vp),
    ("eps", f32), ("silu", i32), ("out", vp), ("out_ld", i32), ("halo", i32), ("H", i32), ("W", i32),
    ("mode", i32), ("sums", vp), ("count", C.c_double), ("n_peers", i32), ("sums_peers", C.c_void_p * 8)])

PtLayerNormArgs = _st("PtLayerNormArgs", [
    ("x", vp), ("ld", i32), ("gamma", vp), ("beta", vp), ("eps", f32), ("out", vp), ("out_ld", i32),
    ("rows", i32), ("C", i32), ("addvec", vp), ("hw", i32), ("F", i32), ("sum_out", vp)])

PtAttnSpatialArgs = _st("PtAttnSpatialArgs", [
    ("tmap_qkv", vp), ("out", vp), ("out_ld", i32), ("S", i32), ("heads", i32), ("C", i32), ("n_img", i32), ("lse", vp)])

PtAttnTemporalArgs = _st("PtAttnTemporalArgs", [
    ("qkv", vp), ("ld", i32), ("out", vp), ("out_ld", i32),
    ("B", i32), ("F", i32), ("HW", i32), ("heads", i32), ("C", i32)])

PtSmallLinearArgs = _st("PtSmallLinearArgs", [
    ("in_", vp), ("in_ld", i32), ("w", vp), ("w_ld", i32), ("bias", vp), ("out", vp), ("out_ld", i32),
    ("M", i32), ("N", i32), ("K", i32), ("act_in_silu", i32), ("act_out_silu", i32), ("accumulate", i32)])

PtSinCosArgs = _st("PtSinCosArgs", [
    ("t", vp), ("sigmas", vp), ("step_index", vp), ("out", vp), ("out_ld", i32), ("M", i32), ("dim", i32)])

PtUpsampleArgs = _st("PtUpsampleArgs", [
    ("x", vp), ("ld", i32), ("out", vp), ("out_ld", i32),
    ("n", i32), ("H", i32), ("W", i32), ("C", i32), ("halo", i32), ("scale", i32)])

PtConvDirectArgs = _st("PtConvDirectArgs", [
    ("x", vp), ("in_nchw_f32", i32), ("in_ld", i32), ("w", vp), ("bias", vp), ("out", vp),
    ("out_ld", i32), ("out_halo", i32),
    ("n", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32), ("stride", i32), ("silu", i32)])

PtLayoutArgs = _st("PtLayoutArgs", [
    ("nchw_const_unused", vp), ("nchw", vp), ("tokens", vp),
    ("n", i32), ("C", i32), ("H", i32), ("W", i32), ("ld", i32), ("halo", i32), ("nchw_f32", i32)])


PtRowBlockCopyArgs = _st("PtRowBlockCopyArgs", [
    ("src", vp), ("dst", vp), ("src_ld", i32), ("dst_ld", i32), ("cols", i32),
    ("src_row", vp), ("dst_row", vp), ("rows", vp), ("n_blocks", i32)])

PtAxpyArgs = _st("PtAxpyArgs", [
    ("x", vp), ("y", vp), ("out", vp), ("ld_x", i32), ("ld_y", i32), ("ld_out", i32), ("rows", i32), ("cols", i32),
    ("scale", f32)])

PtSoftmaxArgs = _st("PtSoftmaxArgs", [
    ("in_", vp), ("out", vp), ("rows", i32), ("cols", i32), ("ld_in", i32), ("ld_out", i32)])

PtTimeConvArgs = _st("PtTimeConvArgs", [
    ("in_", vp), ("ld", i32), ("w", vp), ("bias", vp), ("out", vp), ("B", i32), ("F", i32), ("HW", i32), ("C", i32)])

PtBlurArgs = _st("PtBlurArgs", [
    ("in_", vp), ("out", vp), ("w", vp), ("planes", i32), ("H", i32), ("W", i32), ("k", i32), ("axis", i32)])

PtBicubicArgs = _st("PtBicubicArgs", [
    ("in_", vp), ("C", i32), ("H", i32), ("W", i32), ("S", i32), ("P", i32), ("out_f32", vp), ("out_patches", vp),
    ("ld", i32)])

PtAttnSmallArgs = _st("PtAttnSmallArgs", [
    ("qkv", vp), ("ld", i32), ("out", vp), ("out_ld", i32), ("S", i32), ("heads", i32), ("head_dim", i32)])

PtRasterArgs = _st("PtRasterArgs", [
    ("tracks", vp), ("K", i32), ("F", i32), ("H", i32), ("W", i32), ("order", vp), ("out_f32", vp), ("out_u8", vp),
    ("swap_per_track", i32)])

i64, f64 = C.c_int64, C.c_double

PtEdmLossArgs = _st("PtEdmLossArgs", [
    ("pred", vp), ("pred_ld", i32), ("noisy", vp), ("target", vp), ("sample_stride", i64), ("frame_stride", i64),
    ("sigmas", vp), ("B", i32), ("F", i32), ("C", i32), ("HW", i32), ("weight", f32), ("dpred", vp), ("dpred_ld", i32),
    ("workspace", vp), ("loss", vp), ("accumulate", i32)])

PtGroupNormBwdArgs = _st("PtGroupNormBwdArgs", [
    ("x0", vp), ("x1", vp), ("c0", i32), ("c1", i32), ("ld0", i32), ("ld1", i32), ("dout", vp), ("dout_ld", i32),
    ("halo", i32), ("H", i32), ("W", i32), ("gamma", vp), ("beta", vp), ("eps", f32), ("silu", i32),
    ("rows_per_stat", i32), ("num_stat", i32), ("dx0", vp), ("dx1", vp), ("dld0", i32), ("dld1", i32),
    ("workspace", vp), ("dgb_out", vp), ("accumulate_dgb", i32)])

PtLayerNormBwdArgs = _st("PtLayerNormBwdArgs", [
    ("x", vp), ("ld", i32), ("dout", vp), ("dout_ld", i32), ("gamma", vp), ("eps", f32), ("rows", i32), ("C", i32),
    ("addvec", vp), ("hw", i32), ("F", i32), ("dx", vp), ("dx_ld", i32), ("accumulate_dx", i32), ("partials", vp),
    ("n_blocks", i32), ("dgb_out", vp), ("accumulate_dgb", i32)])

PtColsumArgs = _st("PtColsumArgs", [
    ("x", vp), ("ld", i32), ("halo", i32), ("H", i32), ("W", i32), ("rows_per_group", i64), ("groups", i32), ("C", i32),
    ("scale", f32), ("out", vp), ("accumulate", i32)])

PtAdamWArgs = _st("PtAdamWArgs", [
    ("master", vp), ("grad", vp), ("m", vp), ("v", vp), ("work", vp), ("n", i64), ("lr", f32), ("beta1", f32),
    ("beta2", f32), ("eps", f32), ("weight_decay", f32), ("grad_scale", f32), ("step", i32)])

PtWgradArgs = _st("PtWgradArgs", [
    ("tmap_dt", vp), ("tmap_a", vp), ("rows", i32), ("N", i32), ("K", i32), ("num_taps", i32), ("tap_shift", i32 * 9),
    ("splits", i32), ("partials", vp)])

PtAttnSpatialBwdArgs = _st("PtAttnSpatialBwdArgs", [
    ("qkv", vp), ("ld", i32), ("dout", vp), ("dout_ld", i32), ("lse", vp), ("delta", vp), ("dqkv", vp), ("dld", i32),
    ("S", i32), ("heads", i32), ("C", i32), ("n_img", i32), ("tmap_qkv", vp), ("tmap_dout", vp)])

PtAttnTemporalBwdArgs = _st("PtAttnTemporalBwdArgs", [
    ("qkv", vp), ("ld", i32), ("dout", vp), ("dout_ld", i32), ("dqkv", vp), ("dld", i32),
    ("B", i32), ("F", i32), ("HW", i32), ("heads", i32), ("C", i32)])

PtSmallLinearBwdArgs = _st("PtSmallLinearBwdArgs", [
    ("x", vp), ("x_ld", i32), ("w", vp), ("w_ld", i32), ("dy", vp), ("dy_ld", i32), ("M", i32), ("N", i32), ("K", i32),
    ("act_in_silu", i32), ("dx", vp), ("dx_ld", i32), ("accumulate_dx", i32), ("dw", vp), ("db", vp), ("accumulate_w", i32),
    ("dx_workspace", vp)])

PtColsumGroupedArgs = _st("PtColsumGroupedArgs", [
    ("x", vp), ("ld", i32), ("rows", i64), ("C", i32), ("groups", i32), ("mode", i32), ("ga", i32), ("gb", i32), ("gc", i32),
    ("scale", f32), ("out", vp), ("out_ld", i32), ("accumulate", i32), ("workspace", vp)])

PT_DT_BF16 = 0
PT_DT_F32 = 1

# every symbol include/posetraj_b200.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "pt_last_error": (C.c_char_p, []),
    "pt_version": (C.c_int, []),
    "pt_pdl": (C.c_int, []),
    "pt_launch_count": (C.c_int64, []),
    "pt_sizeof": (C.c_int, [C.c_char_p]),
    "pt_tensormap_encode_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "pt_cfg_euler_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_step_advance": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_gemm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_mlp_geglu": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_gemm_set_trace": (None, [C.c_void_p]),
    "pt_groupnorm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_groupnorm_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "pt_layernorm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_attention_spatial": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_attention_temporal": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_small_linear": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_timestep_sincos": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_upsample2x": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_conv3x3_direct": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_nchw_to_tokens": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_tokens_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_rasterize_tracks": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_row_block_copy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_axpy_bf16": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_rasterize_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "pt_softmax_rows": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_blur_reflect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_bicubic_resize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_attention_small": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_time_conv3": (C.c_int, [C.c_void_p, C.c_void_p]),
    # training step
    "pt_edm_loss": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_edm_loss_workspace_bytes": (C.c_int64, []),
    "pt_groupnorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_groupnorm_bwd_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "pt_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_geglu_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "pt_geglu_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                               C.c_void_p]),
    "pt_colsum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_reduce_partials": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_int32, C.c_void_p]),
    "pt_dot_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_int32,
                              C.c_void_p, C.c_void_p]),
    "pt_transpose_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pt_adamw": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_wgrad": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_attention_delta": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "pt_attention_spatial_bwd": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_attention_temporal_bwd": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_upsample2x_bwd": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_dilate2x": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                              C.c_void_p]),
    "pt_zero_halo": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pt_silu_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "pt_silu_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "pt_small_linear_bwd": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_small_linear_bwd_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "pt_colsum_grouped": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_colsum_grouped_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
}

_lib = None


class PoseTrajLibError(RuntimeError):
    pass


def exported_symbols() -> list[str]:
    return list(_SIGNATURES)


def lib() -> C.CDLL:
    """Load the CUDA library (once). Raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise PoseTrajLibError(
                f"{LIB_PATH} is missing: build it with `python -m posetraj_b200.build` "
                "(there is no fallback path)")
        handle = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().pt_last_error().decode(errors="replace")
        if code == 1:
            raise ValueError(f"{what}: {msg}")
        raise PoseTrajLibError(f"{what}: CUDA error {code}: {msg}")


def aligned_tensormap() -> PtTensorMap:
    """A 64-byte aligned PtTensorMap (ctypes only guarantees 8)."""
    raw = (C.c_uint8 * (128 + 64))()
    addr = C.addressof(raw)
    off = (-addr) % 64
    tm = PtTensorMap.from_buffer(raw, off)
    tm._keepalive = raw
    return tm


def encode_tensormap(base_ptr: int, dims: list[int], strides_bytes: list[int], box: list[int]) -> PtTensorMap:
    rank = len(dims)
    tm = aligned_tensormap()
    d = (C.c_uint64 * rank)(*dims)
    s = (C.c_uint64 * max(1, rank - 1))(*strides_bytes)
    b = (C.c_uint32 * rank)(*box)
    check(lib().pt_tensormap_encode_bf16(C.addressof(tm), C.c_void_p(base_ptr), rank, d, s, b),
          "pt_tensormap_encode_bf16")
    return tm
