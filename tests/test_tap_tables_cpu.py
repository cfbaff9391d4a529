"""CPU check of the implicit-GEMM convolution contract of `pt_gemm` (include/posetraj_b200.h: tap shifts over the
zero-haloed row space, map_mode 1, ostride): a few lines of numpy emulate what the kernel computes from the HOST-side
tap tables and weight layout, and the result must equal torch's conv2d — for the UNet's padding-1 convs
(`ops.conv3x3_taps`, stride 1 and the stride-2 Downsample2D) and for the VAE encoder's `Downsample2D(padding=0)`
(pad right/bottom, taps (0..2, 0..2); posetraj_b200/vae.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def haloed_rows(x):
    """[n, C, H, W] -> zero-haloed token rows [n*(H+1)*(W+1), C] (one extra zero column / row per image)."""
    n, C, H, W = x.shape
    out = torch.zeros(n, H + 1, W + 1, C)
    out[:, :H, :W] = x.permute(0, 2, 3, 1)
    return out.reshape(n * (H + 1) * (W + 1), C).numpy()


def emulate_gemm(rows, w_kmajor, taps, n, H, W, ostride):
    """D[r] = sum_t A[r + shift_t] @ W_t^T with zero fill outside the tensor (TMA OOB), then map_mode 1: keep rows with
    y < H, x < W, y % s == x % s == 0 -> compact output [n, H/s, W/s, N]."""
    R, C = rows.shape
    N = w_kmajor.shape[0]
    acc = np.zeros((R, N), dtype=np.float64)
    for t, shift in enumerate(taps):
        a = np.zeros_like(rows, dtype=np.float64)
        lo, hi = max(0, -shift), min(R, R - shift)
        a[lo:hi] = rows[lo + shift:hi + shift]
        acc += a @ w_kmajor[:, t * C:(t + 1) * C].T.astype(np.float64)
    acc = acc.reshape(n, H + 1, W + 1, N)[:, :H:ostride, :W:ostride]
    return torch.from_numpy(acc).permute(0, 3, 1, 2).float()


def kmajor(weight):
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin], K index = (ky*3 + kx)*Cin + ci — through the product's own WeightStore.conv3
    (bf16 storage: the weights of these tests are bf16-representable)."""
    from posetraj_b200.engine import WeightStore
    return WeightStore({"w": weight}, "cpu").conv3("w").float().numpy()


@pytest.mark.parametrize("stride", [1, 2])
def test_padding1_taps_match_conv2d(stride):
    from posetraj_b200.ops import conv3x3_taps
    g = torch.Generator().manual_seed(0)
    x, w = torch.randn(2, 5, 6, 8, generator=g), torch.randn(7, 5, 3, 3, generator=g).bfloat16().float()
    got = emulate_gemm(haloed_rows(x), kmajor(w), conv3x3_taps(8), 2, 6, 8, stride)
    want = F.conv2d(x, w, stride=stride, padding=1)
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-4)


def test_vae_downsample_taps_match_pad_right_bottom_conv():
    g = torch.Generator().manual_seed(1)
    x, w = torch.randn(2, 4, 6, 8, generator=g), torch.randn(3, 4, 3, 3, generator=g).bfloat16().float()
    taps = [ky * (8 + 1) + kx for ky in range(3) for kx in range(3)]      # posetraj_b200/vae.py VaeEncodePlan
    got = emulate_gemm(haloed_rows(x), kmajor(w), taps, 2, 6, 8, 2)
    want = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, stride=2, padding=0)
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-4)


def test_temporal_taps_match_conv3d():
    """TemporalResnetBlock's (3,1,1) conv: taps (-HW, 0, +HW) over rows (b, f, hw), zero fill at the ends of each
    BATCH (the A operand is a rank-3 tensor map, rows of another batch are out of bounds)."""
    g = torch.Generator().manual_seed(2)
    B, Fr, C, HW, N = 2, 4, 3, 5, 6
    x, w = torch.randn(B, C, Fr, HW, 1, generator=g), torch.randn(N, C, 3, 1, 1, generator=g).bfloat16().float()
    rows = x[..., 0].permute(0, 2, 3, 1).reshape(B, Fr * HW, C).numpy()
    from posetraj_b200.engine import WeightStore
    wk = WeightStore({"w": w}, "cpu").tconv("w").float().numpy()                 # K = kt*C + ci
    out = np.zeros((B, Fr * HW, N))
    for b in range(B):
        for t, shift in enumerate((-HW, 0, HW)):
            a = np.zeros_like(rows[b])
            lo, hi = max(0, -shift), min(Fr * HW, Fr * HW - shift)
            a[lo:hi] = rows[b][lo + shift:hi + shift]
            out[b] += a @ wk[:, t * C:(t + 1) * C].T
    want = F.conv3d(x, w, padding=(1, 0, 0))[..., 0].permute(0, 2, 3, 1).reshape(B, Fr * HW, N)
    assert torch.allclose(torch.from_numpy(out).float(), want, atol=1e-4)
