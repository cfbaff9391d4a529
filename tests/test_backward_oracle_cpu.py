"""The explicit backward formulas the next round's kernels will implement (oracle/backward.py) against torch autograd,
in the forward library's data layouts (zero-haloed row space, token-major rows, K-major per-tap weights)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from test_tap_tables_cpu import haloed_rows, kmajor


def _unhalo(rows, n, H, W):
    C = rows.shape[1]
    return torch.from_numpy(rows.reshape(n, H + 1, W + 1, C)[:, :H, :W]).permute(0, 3, 1, 2)


@pytest.mark.parametrize("stride", [1, 2])
def test_conv_dgrad_wgrad_in_the_haloed_row_space(stride):
    from oracle.backward import conv_rows_dgrad, conv_rows_forward, conv_rows_wgrad, valid_mask
    from posetraj_b200.ops import conv3x3_taps
    g = torch.Generator().manual_seed(0)
    n, Cin, Cout, H, W = 2, 5, 7, 6, 8
    x = torch.randn(n, Cin, H, W, generator=g, requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, generator=g).bfloat16().float().requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    taps = conv3x3_taps(W)
    a = haloed_rows(x.detach()).astype(np.float64)
    wk = kmajor(w.detach()).astype(np.float64)
    # forward in row space reproduces conv2d on the valid rows
    d = conv_rows_forward(a, wk, taps).reshape(n, H + 1, W + 1, Cout)[:, :H:stride, :W:stride]
    assert np.allclose(d, y.detach().permute(0, 2, 3, 1).numpy(), atol=1e-4)
    # dD scattered onto the haloed row space (zero on halo / skipped rows)
    dD = np.zeros((n, H + 1, W + 1, Cout))
    dD[:, :H:stride, :W:stride] = dy.permute(0, 2, 3, 1).numpy()
    dD = dD.reshape(-1, Cout)
    assert np.allclose(dD * valid_mask(n, H, W, stride), dD)
    dA = conv_rows_dgrad(dD, wk, taps, Cin)
    assert torch.allclose(_unhalo(dA, n, H, W).float(), x.grad, atol=1e-4)
    dW = conv_rows_wgrad(dD, a, taps)                       # [Cout, 9*Cin], K index = (ky*3+kx)*Cin + ci
    want = w.grad.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).numpy()
    assert np.allclose(dW, want, atol=1e-3)


@pytest.mark.parametrize("rows_per_stat,use_silu", [(12, True), (36, True), (12, False)])
def test_groupnorm_silu_backward(rows_per_stat, use_silu):
    from oracle.backward import groupnorm_silu_backward
    g = torch.Generator().manual_seed(1)
    n_stat, C, eps = 3, 64, 1e-6
    rows = n_stat * rows_per_stat
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(C, generator=g)).requires_grad_(True)
    # torch reference: [n_stat, C, rows_per_stat] so that every statistics block is one "sample"
    xr = x.reshape(n_stat, rows_per_stat, C).permute(0, 2, 1)
    y = F.group_norm(xr, 32, gamma, beta, eps)
    if use_silu:
        y = F.silu(y)
    d_out = torch.randn(rows, C, generator=g)
    y.backward(d_out.reshape(n_stat, rows_per_stat, C).permute(0, 2, 1))
    dx, dgam, dbet = groupnorm_silu_backward(x.detach(), gamma.detach(), beta.detach(), d_out, rows_per_stat, eps, use_silu)
    assert torch.allclose(dx.float(), x.grad, atol=2e-5)
    assert torch.allclose(dgam.float(), gamma.grad, atol=2e-4) and torch.allclose(dbet.float(), beta.grad, atol=2e-4)


def test_layernorm_backward():
    from oracle.backward import layernorm_backward
    g = torch.Generator().manual_seed(2)
    x = torch.randn(40, 96, generator=g, requires_grad=True)
    gamma = (1 + 0.2 * torch.randn(96, generator=g)).requires_grad_(True)
    beta = torch.zeros(96, requires_grad=True)
    d_out = torch.randn(40, 96, generator=g)
    F.layer_norm(x, (96,), gamma, beta, 1e-5).backward(d_out)
    dx, dgam, dbet = layernorm_backward(x.detach(), gamma.detach(), d_out)
    assert torch.allclose(dx.float(), x.grad, atol=2e-5)
    assert torch.allclose(dgam.float(), gamma.grad, atol=2e-4) and torch.allclose(dbet.float(), beta.grad, atol=2e-4)


def test_geglu_backward():
    from oracle.backward import geglu_backward
    g = torch.Generator().manual_seed(3)
    v = torch.randn(50, 32, generator=g, requires_grad=True)
    gt = (torch.randn(50, 32, generator=g) * 2).requires_grad_(True)
    d_out = torch.randn(50, 32, generator=g)
    (v * F.gelu(gt)).backward(d_out)
    dv, dg = geglu_backward(v.detach(), gt.detach(), d_out)
    assert torch.allclose(dv.float(), v.grad, atol=1e-5) and torch.allclose(dg.float(), gt.grad, atol=1e-5)


def test_attention_backward_flash_formulation():
    from oracle.backward import attention_backward
    g = torch.Generator().manual_seed(4)
    S, hd = 45, 64
    q, k, v = (torch.randn(S, hd, generator=g, requires_grad=True) for _ in range(3))
    d_out = torch.randn(S, hd, generator=g)
    F.scaled_dot_product_attention(q[None, None], k[None, None], v[None, None])[0, 0].backward(d_out)
    dq, dk, dv = attention_backward(q.detach(), k.detach(), v.detach(), d_out, 1.0 / math.sqrt(hd))
    for got, want in ((dq, q.grad), (dk, k.grad), (dv, v.grad)):
        assert torch.allclose(got.float(), want, atol=2e-5)
