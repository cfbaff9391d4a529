"""ORACLE (test infrastructure, never the product path) — explicit backward formulas of the hot-path operators, written
out the way the sm_100a kernels of the next round will compute them (config-4 backward), and
checked against torch autograd in tests/test_backward_oracle_cpu.py.  Nothing here is taken from the reference (it
relies on autograd, scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1470); these are the specifications the backward
kernels will be held to, in the data layouts of the forward library:

  * implicit-GEMM convolution in the zero-haloed row space (include/posetraj_b200.h, PtGemmArgs):
        forward   D[r]  = sum_t A[r + s_t] W_t^T
        dgrad     dA[r] = sum_t dD[r - s_t] W_t          (the same kernel with negated shifts and W_t as [K, N])
        wgrad     dW_t  = sum_r dD[r]^T A[r + s_t]       (one [N, K] GEMM per tap, reduction over rows)
    with dD zero on the halo / invalid rows and dA discarded there;
  * GroupNorm(32)+SiLU, LayerNorm, GEGLU (exact erf), softmax attention (the flash formulation: no P stored).

Only tests/ may import this.
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------------------------
# implicit-GEMM convolution in the haloed row space
# ---------------------------------------------------------------------------------------------------------------
def shift_rows(a: np.ndarray, shift: int) -> np.ndarray:
    """B[r] = A[r + shift], zero outside the tensor (TMA out-of-bounds fill)."""
    out = np.zeros_like(a)
    R = a.shape[0]
    lo, hi = max(0, -shift), min(R, R - shift)
    if hi > lo:
        out[lo:hi] = a[lo + shift:hi + shift]
    return out


def conv_rows_forward(a, w_kmajor, taps):
    K = a.shape[1]
    return sum(shift_rows(a, s) @ w_kmajor[:, t * K:(t + 1) * K].T for t, s in enumerate(taps))


def conv_rows_dgrad(d_out, w_kmajor, taps, K):
    """dA[r] = sum_t dD[r - s_t] W_t — d_out must already be zero on rows that are not real outputs."""
    return sum(shift_rows(d_out, -s) @ w_kmajor[:, t * K:(t + 1) * K] for t, s in enumerate(taps))


def conv_rows_wgrad(d_out, a, taps):
    """dW[:, t*K:(t+1)*K] = dD^T (A shifted by s_t)."""
    return np.concatenate([d_out.T @ shift_rows(a, s) for s in taps], axis=1)


def valid_mask(n, H, W, ostride=1):
    """1 on the haloed rows that are real (strided) output pixels (map_mode 1 of the forward epilogue)."""
    m = np.zeros((n, H + 1, W + 1), dtype=np.float64)
    m[:, :H:ostride, :W:ostride] = 1.0
    return m.reshape(-1, 1)


# ---------------------------------------------------------------------------------------------------------------
# normalisations
# ---------------------------------------------------------------------------------------------------------------
def silu(x):
    return x / (1.0 + torch.exp(-x))


def silu_grad(x):
    s = torch.sigmoid(x)
    return s * (1.0 + x * (1.0 - s))


def groupnorm_silu_backward(x, gamma, beta, d_out, rows_per_stat: int, eps: float, use_silu: bool = True, groups: int = 32):
    """x, d_out: token-major [rows, C]; statistics per (block of rows_per_stat rows, group of C/groups channels).
    Returns (dx, dgamma, dbeta).  With xh = (x - mu) rstd, y = gamma xh + beta, out = silu(y):
        g = d_out silu'(y) gamma ;  dx = rstd (g - mean(g) - xh mean(g xh))   (means over the statistics group)."""
    rows, C = x.shape
    S, cg = rows // rows_per_stat, C // groups
    xs = x.reshape(S, rows_per_stat, groups, cg).double()
    mu = xs.mean(dim=(1, 3), keepdim=True)
    var = xs.var(dim=(1, 3), unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    xh = (xs - mu) * rstd
    gam = gamma.double().reshape(1, 1, groups, cg)
    y = xh * gam + beta.double().reshape(1, 1, groups, cg)
    dy = d_out.reshape(S, rows_per_stat, groups, cg).double()
    if use_silu:
        dy = dy * silu_grad(y)
    g = dy * gam
    dx = rstd * (g - g.mean(dim=(1, 3), keepdim=True) - xh * (g * xh).mean(dim=(1, 3), keepdim=True))
    dgamma = (dy * xh).sum(dim=(0, 1)).reshape(C)
    dbeta = dy.sum(dim=(0, 1)).reshape(C)
    return dx.reshape(rows, C), dgamma, dbeta


def layernorm_backward(x, gamma, d_out, eps: float = 1e-5):
    """Row-wise: dx = rstd (g - mean(g) - xh mean(g xh)), g = d_out gamma."""
    x = x.double()
    mu = x.mean(-1, keepdim=True)
    rstd = (x.var(-1, unbiased=False, keepdim=True) + eps).rsqrt()
    xh = (x - mu) * rstd
    g = d_out.double() * gamma.double()
    dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    return dx, (d_out.double() * xh).sum(0), d_out.double().sum(0)


# ---------------------------------------------------------------------------------------------------------------
# GEGLU and attention
# ---------------------------------------------------------------------------------------------------------------
def geglu_backward(value, gate, d_out):
    """out = value * gelu(gate), exact erf: gelu'(g) = Phi(g) + g phi(g)."""
    value, gate, d_out = value.double(), gate.double(), d_out.double()
    Phi = 0.5 * (1.0 + torch.erf(gate / math.sqrt(2.0)))
    phi = torch.exp(-0.5 * gate * gate) / math.sqrt(2.0 * math.pi)
    return d_out * gate * Phi, d_out * value * (Phi + gate * phi)


def attention_backward(q, k, v, d_out, scale: float):
    """Flash formulation for one head: P = softmax(scale q k^T), O = P v.
        D = rowsum(dO * O) ; dV = P^T dO ; dP = dO v^T ; dS = P * (dP - D) ; dQ = scale dS k ; dK = scale dS^T q."""
    q, k, v, d_out = q.double(), k.double(), v.double(), d_out.double()
    P = torch.softmax(scale * q @ k.T, dim=-1)
    O = P @ v
    D = (d_out * O).sum(-1, keepdim=True)
    dV = P.T @ d_out
    dS = P * (d_out @ v.T - D)
    return scale * dS @ k, scale * dS.T @ q, dV
