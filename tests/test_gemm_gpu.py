"""Kernel-level numerics of the tcgen05 GEMM / implicit-GEMM conv against plain torch fp32 (GPU tests).

Inputs are bf16-rounded so the only differences are accumulation order and the final bf16 rounding:
tolerance rel-L2 <= 4e-3 (bf16 output rounding is ~2e-3 rms).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 4e-3


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def sp():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, scale=1.0, dev="cuda"):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 128, 128, None), (1000, 320, 320, 160),
                                      (5000, 1280, 640, 256), (333, 96, 192, 96), (80000, 320, 320, None), (4096, 640, 2560, None),
                                      (1260, 1280, 1280, None)])
@pytest.mark.parametrize("pair", [False, True])
def test_linear(cuda_dev, pair, M, N, K, bn):
    from posetraj_b200.ops import Gemm
    torch.manual_seed(0)
    a = rnd(M, K)
    w = rnd(N, K, scale=1 / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    Gemm(a, w, out, bias=bias, block_n=bn, cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    assert rel_l2(out, ref) < TOL


def test_linear_f32_out_small_n(cuda_dev):
    from posetraj_b200.ops import Gemm
    torch.manual_seed(1)
    M, N, K = 777, 4, 320
    a = rnd(M, K)
    w = rnd(N, K, scale=1 / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    Gemm(a, w, out, bias=bias, block_n=32).launch(sp())
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    assert rel_l2(out, ref) < 1e-4


@pytest.mark.parametrize("pair", [False, True])
def test_epilogue_residual_rowvec_out2(cuda_dev, pair):
    from posetraj_b200.ops import Gemm
    torch.manual_seed(2)
    groups, per = 4, 300
    M, N, K = groups * per, 320, 640
    a = rnd(M, K)
    w = rnd(N, K, scale=1 / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    rowvec = torch.randn(groups, N, device="cuda")
    r1, r2, aux = rnd(M, N), rnd(M, N), rnd(M, N)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty_like(out)
    Gemm(a, w, out, bias=bias, rowvec=rowvec, rowvec_mode=1, rv=(per, 1, 1), acc_scale=0.375,
         res1=r1, res1_scale=1.0, res2=r2, res2_scale=0.625, out2=out2, aux=aux, aux_scale=3.0, cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    core = a.float() @ w.float().t() + bias + rowvec.repeat_interleave(per, 0)
    ref = 0.375 * core + r1.float() + 0.625 * r2.float()
    assert rel_l2(out, ref) < TOL
    assert rel_l2(out2, ref + 3.0 * aux.float()) < TOL


@pytest.mark.parametrize("pair", [False, True])
def test_rowvec_mode2(cuda_dev, pair):
    """Temporal cross-attention quirk: vector index ((row // (F*HW)) * HW + row % HW) % B."""
    from posetraj_b200.ops import Gemm
    torch.manual_seed(3)
    B, Fr, HW, N, K = 2, 3, 45, 64, 64
    M = B * Fr * HW
    a = rnd(M, K)
    w = rnd(N, K, scale=1 / math.sqrt(K))
    rowvec = torch.randn(B, N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    Gemm(a, w, out, rowvec=rowvec, rowvec_mode=2, rv=(Fr * HW, HW, B), cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    rows = torch.arange(M, device="cuda")
    g = ((rows // (Fr * HW)) * HW + rows % HW) % B
    ref = a.float() @ w.float().t() + rowvec[g]
    assert rel_l2(out, ref) < TOL


@pytest.mark.parametrize("M,C,bn", [(512, 320, None), (2000, 640, 256), (700, 64, 64)])
@pytest.mark.parametrize("pair", [False, True])
def test_geglu(cuda_dev, pair, M, C, bn):
    from posetraj_b200.ops import Gemm
    torch.manual_seed(4)
    a = rnd(M, C)
    w = rnd(8 * C, C, scale=1 / math.sqrt(C))
    bias = torch.randn(8 * C, device="cuda")
    out = torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16)
    Gemm(a, w, out, bias=bias, geglu=True, block_n=bn, cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    h = a.float() @ w.float().t() + bias
    x, g = h.chunk(2, dim=-1)
    ref = x * F.gelu(g)
    assert rel_l2(out, ref) < TOL


def test_geglu_gate_pointwise_incl_negative_tail(cuda_dev):
    """The epilogue's gate is value * g * Phi(g) as hv + hv tanh(u) with ONE tanh.approx (common.cuh geglu_gate_tanh).  With a
    one-hot weight matrix the GEMM is a copy, so every output is the gate of a known (value, g) pair: gates from -9 to +9,
    element-wise bound = bf16 rounding of the result + |value g| * 3e-4 (tanh.approx is good to 2^-11 relative; the
    cancellation 1 + tanh(u) in the negative tail turns that into an absolute error of that size)."""
    from posetraj_b200.ops import Gemm
    torch.manual_seed(14)
    M, C = 1024, 128
    a = torch.zeros(M, C, device="cuda")
    a[:, :64] = torch.randn(M, 64, device="cuda") * 2.0                       # values
    a[:, 64:] = torch.linspace(-9.0, 9.0, M * 64, device="cuda").view(M, 64)   # gates
    a = a.to(torch.bfloat16)
    w = torch.eye(C, device="cuda")                 # rows 0..63 -> x (values), rows 64..127 -> g (gates)
    out = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
    Gemm(a, w.to(torch.bfloat16), out, geglu=True).launch(sp())
    torch.cuda.synchronize()
    x, g = a.float()[:, :64], a.float()[:, 64:]
    ref = x * F.gelu(g)
    err = (out.float() - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + (x * g).abs() * 3e-4 + 1e-6
    assert bool((err <= bound).all()), float((err - bound).max())


@pytest.mark.parametrize("pair", [False, True])
def test_two_k_sources(cuda_dev, pair):
    from posetraj_b200.ops import Gemm
    torch.manual_seed(5)
    M, K0, K1, N = 900, 640, 320, 640
    a0, a1 = rnd(M, K0), rnd(M, K1)
    w = rnd(N, K0 + K1, scale=1 / math.sqrt(K0 + K1))
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    Gemm(a0, w, out, a1=a1, cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    ref = torch.cat([a0, a1], 1).float() @ w.float().t()
    assert rel_l2(out, ref) < TOL


def halo_pack(x_nhwc):
    """[n,H,W,C] -> zero-haloed rows [n*(H+1)*(W+1), C] (one zero column right, one zero row below)."""
    n, H, W, Cc = x_nhwc.shape
    p = torch.zeros(n, H + 1, W + 1, Cc, device=x_nhwc.device, dtype=x_nhwc.dtype)
    p[:, :H, :W] = x_nhwc
    return p.reshape(n * (H + 1) * (W + 1), Cc)


@pytest.mark.parametrize("n,H,W,Cin,Cout,stride", [(3, 10, 18, 64, 64, 1), (4, 20, 36, 320, 640, 1),
                                                   (5, 5, 9, 128, 192, 1), (3, 10, 18, 128, 128, 2),
                                                   (2, 40, 72, 64, 32, 1)])
@pytest.mark.parametrize("pair", [False, True])
def test_conv3x3(cuda_dev, pair, n, H, W, Cin, Cout, stride):
    from posetraj_b200.ops import Gemm, conv3x3_taps
    torch.manual_seed(6)
    x = rnd(n, H, W, Cin)
    wt = rnd(Cout, Cin, 3, 3, scale=1 / math.sqrt(9 * Cin))
    bias = torch.randn(Cout, device="cuda")
    a = halo_pack(x)
    w2 = wt.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    oH, oW = (H + stride - 1) // stride, (W + stride - 1) // stride
    out = torch.zeros(n * oH * oW, Cout, device="cuda", dtype=torch.bfloat16)
    Gemm(a, w2, out, taps=conv3x3_taps(W), bias=bias, halo=(H, W), ostride=stride, cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=stride, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(n * oH * oW, Cout)
    assert rel_l2(out, ref) < TOL


@pytest.mark.parametrize("B,Fr,HW,Cc", [(2, 14, 45, 128), (2, 5, 180, 320), (1, 3, 720, 64)])
@pytest.mark.parametrize("pair", [False, True])
def test_temporal_conv(cuda_dev, pair, B, Fr, HW, Cc):
    from posetraj_b200.ops import Gemm
    torch.manual_seed(7)
    x = rnd(B, Fr, HW, Cc)
    wt = rnd(Cc, Cc, 3, scale=1 / math.sqrt(3 * Cc))
    bias = torch.randn(Cc, device="cuda")
    res = rnd(B * Fr * HW, Cc)
    a = x.reshape(B * Fr * HW, Cc)
    w2 = wt.permute(0, 2, 1).reshape(Cc, 3 * Cc).contiguous()
    out = torch.empty(B * Fr * HW, Cc, device="cuda", dtype=torch.bfloat16)
    Gemm(a, w2, out, batches=B, taps=(-HW, 0, HW), bias=bias, res1=res, cta_pair=pair).launch(sp())
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2)  # [B, C, F, HW]
    ref = F.conv2d(xr, wt.float()[..., None], bias, padding=(1, 0))  # kernel (3,1) over (F, HW)
    ref = ref.permute(0, 2, 3, 1).reshape(B * Fr * HW, Cc) + res.float()
    assert rel_l2(out, ref) < TOL


def test_bad_args_raise(cuda_dev):
    from posetraj_b200.ops import Gemm
    a = rnd(128, 64)
    w = rnd(64, 64)
    out = torch.empty(128, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        Gemm(a, w, out, block_n=48).launch(sp())


@pytest.mark.parametrize("M,C", [(256, 64), (300, 64), (1000, 128), (777, 192), (5000, 256), (2880 * 3 + 17, 320), (80640, 320)])
@pytest.mark.parametrize("mix", [False, True])
def test_fused_geglu_mlp(cuda_dev, M, C, mix):
    """ops.FusedMlp (one kernel, hidden activations stay on the SM) against torch fp32 of diffusers' FeedForward:
    net.0 = GEGLU(proj), net.2 = Linear, plus the residual / AlphaBlender terms the engine folds into it.  The hidden
    activations are rounded to bf16 before the second contraction exactly as in the two-kernel path, so the bound is
    the same 4e-3."""
    from posetraj_b200.ops import FusedMlp
    torch.manual_seed(2)
    H = 4 * C
    x = rnd(M, C)
    w1 = rnd(2 * H, C, scale=1 / math.sqrt(C))
    b1 = torch.randn(2 * H, device="cuda") * 0.1
    w2 = rnd(C, H, scale=1 / math.sqrt(H))
    b2 = torch.randn(C, device="cuda") * 0.1
    res1 = rnd(M, C)
    res2 = rnd(M, C) if mix else None
    kw = dict(acc_scale=0.4, res1_scale=0.4, res2_scale=0.6) if mix else {}
    out = torch.zeros(M, C, device="cuda", dtype=torch.bfloat16)
    FusedMlp(x, w1, b1, w2, b2, out, res1=res1, res2=res2, name="test.mlp", **kw).launch(sp())
    torch.cuda.synchronize()
    h = x.float() @ w1.float().t() + b1
    hid = (h[:, :H] * F.gelu(h[:, H:])).to(torch.bfloat16).float()
    ref = hid @ w2.float().t() + b2
    if mix:
        ref = 0.4 * ref + 0.4 * res1.float() + 0.6 * res2.float()
    else:
        ref = ref + res1.float()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < TOL, rel_l2(out, ref)


def test_fused_mlp_matches_the_two_kernel_path(cuda_dev):
    """Same operands through pt_gemm(geglu) + pt_gemm(out): the two lowerings of one feed-forward agree to bf16 rounding."""
    from posetraj_b200.ops import FusedMlp, Gemm
    torch.manual_seed(3)
    M, C = 2880 * 2, 320
    H = 4 * C
    x, w1, w2 = rnd(M, C), rnd(2 * H, C, scale=1 / math.sqrt(C)), rnd(C, H, scale=1 / math.sqrt(H))
    b1, b2 = torch.randn(2 * H, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1
    res = rnd(M, C)
    a = torch.zeros(M, C, device="cuda", dtype=torch.bfloat16)
    FusedMlp(x, w1, b1, w2, b2, a, res1=res).launch(sp())
    hid = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(M, C, device="cuda", dtype=torch.bfloat16)
    Gemm(x, w1, hid, geglu=True, bias=b1).launch(sp())
    Gemm(hid, w2, b, bias=b2, res1=res).launch(sp())
    torch.cuda.synchronize()
    assert rel_l2(a, b) < TOL
