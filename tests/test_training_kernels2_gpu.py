"""The backward kernels added for the whole-network training step (csrc/train_attn.cu, train_misc.cu) against torch
autograd / plain torch fp32 on the same bf16-rounded inputs (BASELINE configs[3])."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def sp():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("n_img,S,heads", [(2, 2880, 5), (3, 720, 2), (2, 180, 3), (4, 45, 2), (1, 200, 1)])
def test_attention_spatial_backward(cuda_dev, n_img, S, heads):
    from posetraj_b200 import ops, training
    torch.manual_seed(0)
    Cc = heads * 64
    qkv = rnd(n_img * S, 3 * Cc, scale=1.5)
    dout = rnd(n_img * S, Cc)
    out = torch.empty(n_img * S, Cc, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(n_img * heads * S, device="cuda", dtype=torch.float32)
    ops.AttnSpatial(qkv, out, n_img=n_img, heads=heads, lse=lse).launch(sp())
    dqkv = training.attention_spatial_backward(qkv, out, dout, lse, n_img=n_img, heads=heads)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    q, k, v = [t.view(n_img, S, heads, 64).transpose(1, 2) for t in x.split(Cc, dim=1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n_img * S, Cc)
    # log-sum-exp of the forward kernel (log2 domain)
    s2 = (q @ k.transpose(-1, -2)) * (0.125 * 1.4426950408889634)
    want_lse = torch.logsumexp(s2 * 0.6931471805599453, -1) / 0.6931471805599453
    assert (lse.view(n_img, heads, S) - want_lse.detach()).abs().max() < 2e-2
    assert rel_l2(out, ref.detach()) < 8e-3
    ref.backward(dout.float())
    g = x.grad
    for i, nm in enumerate("qkv"):
        assert rel_l2(dqkv[:, i * Cc:(i + 1) * Cc], g[:, i * Cc:(i + 1) * Cc]) < 1.2e-2, nm


@pytest.mark.parametrize("B,Fr,HW,heads", [(2, 14, 90, 5), (1, 25, 33, 2), (2, 1, 40, 1), (1, 3, 7, 3)])
def test_attention_temporal_backward(cuda_dev, B, Fr, HW, heads):
    from posetraj_b200 import training
    torch.manual_seed(1)
    Cc = heads * 64
    rows = B * Fr * HW
    qkv = rnd(rows, 3 * Cc, scale=1.5)
    dout = rnd(rows, Cc)
    dqkv = training.attention_temporal_backward(qkv, dout, batch=B, frames=Fr, hw=HW, heads=heads)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    # rows (b*F + f)*HW + s -> [B, HW, heads, F, 64]
    q, k, v = [t.view(B, Fr, HW, heads, 64).permute(0, 2, 3, 1, 4) for t in x.split(Cc, dim=1)]
    o = F.scaled_dot_product_attention(q, k, v)            # [B, HW, heads, F, 64]
    o = o.permute(0, 3, 1, 2, 4).reshape(rows, Cc)
    o.backward(dout.float())
    assert rel_l2(dqkv, x.grad) < 8e-3


def test_upsample_dilate_halo_silu(cuda_dev):
    from posetraj_b200 import ops, training
    torch.manual_seed(2)
    n, H, W, Cc = 3, 5, 7, 64
    # nearest x2 into the zero-haloed layout: backward = sum of the 2x2 children
    g = rnd(n * (2 * H + 1) * (2 * W + 1), Cc)
    dx = training.upsample_backward(g, n=n, H=H, W=W, halo=True, scale=2)
    gi = g.float().view(n, 2 * H + 1, 2 * W + 1, Cc)[:, :2 * H, :2 * W]
    want = gi.reshape(n, H, 2, W, 2, Cc).sum((2, 4)).reshape(n * H * W, Cc)
    assert rel_l2(dx, want) < 4e-3
    # halo copy (scale 1): backward strips the halo
    g1 = rnd(n * (H + 1) * (W + 1), Cc)
    dx1 = training.upsample_backward(g1, n=n, H=H, W=W, halo=True, scale=1)
    assert torch.equal(dx1.view(n, H, W, Cc), g1.view(n, H + 1, W + 1, Cc)[:, :H, :W])
    # dilation of a stride-2 output gradient (compact and haloed sources), odd sizes
    oH, oW = (H + 1) // 2, (W + 1) // 2
    src = rnd(n * oH * oW, Cc)
    d = training.dilate2x(src, n=n, H=H, W=W, src_halo=False).view(n, H + 1, W + 1, Cc)
    want = torch.zeros_like(d)
    want[:, 0:H:2, 0:W:2] = src.view(n, oH, oW, Cc)
    assert torch.equal(d, want)
    srch = torch.zeros(n, oH + 1, oW + 1, Cc, device="cuda", dtype=torch.bfloat16)
    srch[:, :oH, :oW] = src.view(n, oH, oW, Cc)
    d2 = training.dilate2x(srch.view(-1, Cc), n=n, H=H, W=W, src_halo=True).view(n, H + 1, W + 1, Cc)
    assert torch.equal(d2, want)
    # zero_halo
    z = rnd(n * (H + 1) * (W + 1), Cc)
    keep = z.clone().view(n, H + 1, W + 1, Cc)
    training.zero_halo(z, n=n, H=H, W=W)
    zz = z.view(n, H + 1, W + 1, Cc)
    assert zz[:, H].abs().max() == 0 and zz[:, :, W].abs().max() == 0 and torch.equal(zz[:, :H, :W], keep[:, :H, :W])
    # SiLU forward / backward
    x = rnd(1000, 96, scale=3.0)
    dy = rnd(1000, 96)
    y = training.silu_forward(x)
    dxs = training.silu_backward(x, dy)
    xf = x.float().requires_grad_(True)
    yr = F.silu(xf)
    yr.backward(dy.float())
    assert rel_l2(y, yr.detach()) < 4e-3 and rel_l2(dxs, xf.grad) < 4e-3


@pytest.mark.parametrize("M,N,K,act", [(2, 1280, 320, False), (2, 1500, 1280, True), (14, 320, 1280, True), (28, 256, 12, False)])
def test_small_linear_backward(cuda_dev, M, N, K, act):
    from posetraj_b200 import ops, training
    torch.manual_seed(3)
    x = torch.randn(M, K, device="cuda")
    w = rnd(N, K, scale=0.05)
    b = torch.randn(N, device="cuda")
    dy = torch.randn(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.SmallLinear(x, w, out, b, act_in_silu=act).launch(sp())
    dx, dw, db = training.small_linear_backward(x, w, dy, act_in_silu=act)
    torch.cuda.synchronize()
    xf = x.clone().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    bf = b.clone().requires_grad_(True)
    y = F.linear(F.silu(xf) if act else xf, wf, bf)
    assert rel_l2(out, y.detach()) < 1e-4
    y.backward(dy)
    assert rel_l2(dx, xf.grad) < 1e-4 and rel_l2(dw, wf.grad) < 1e-4 and rel_l2(db, bf.grad) < 1e-4
    # accumulate variants
    dx2, dw2, db2 = training.small_linear_backward(x, w, dy, act_in_silu=act, dw=dw.clone(), db=db.clone(), accumulate_w=True,
                                                   dx=dx.clone(), accumulate_dx=True)
    assert rel_l2(dw2, 2 * wf.grad) < 1e-4 and rel_l2(db2, 2 * bf.grad) < 1e-4 and rel_l2(dx2, 2 * xf.grad) < 1e-4


def test_colsum_grouped(cuda_dev):
    from posetraj_b200 import training
    torch.manual_seed(4)
    B, Fr, HW, Cc = 2, 5, 37, 96
    rows = B * Fr * HW
    x = rnd(rows, Cc)
    xf = x.float()
    r = torch.arange(rows, device="cuda")
    # mode 1: contiguous groups (per batch row)
    got = training.colsum_grouped(x, groups=B, mode=1, ga=Fr * HW)
    assert rel_l2(got, xf.view(B, Fr * HW, Cc).sum(1)) < 1e-5
    # mode 2: ((r / (F*HW)) * HW + r % HW) % B — the reference's mis-aligned temporal context broadcast
    g2 = ((r // (Fr * HW)) * HW + r % HW) % B
    want = torch.zeros(B, Cc, device="cuda").index_add_(0, g2, xf)
    got = training.colsum_grouped(x, groups=B, mode=2, ga=Fr * HW, gb=HW, gc=B, scale=0.5)
    assert rel_l2(got, 0.5 * want) < 1e-5
    # mode 3: frame index
    g3 = (r // HW) % Fr
    want = torch.zeros(Fr, Cc, device="cuda").index_add_(0, g3, xf)
    got = training.colsum_grouped(x, groups=Fr, mode=3, ga=HW, gc=Fr)
    assert rel_l2(got, want) < 1e-5
    acc = training.colsum_grouped(x, groups=Fr, mode=3, ga=HW, gc=Fr, out=got.clone(), accumulate=True)
    assert rel_l2(acc, 2 * want) < 1e-5
    # a big one (bias gradient of a level-0 layer), strided output view
    xb = rnd(80640, 320)
    out = torch.zeros(1, 640, device="cuda")
    training.colsum_grouped(xb, groups=1, mode=1, ga=80640, out=out[:, 320:], accumulate=False)
    assert rel_l2(out[0, 320:], xb.float().sum(0)) < 1e-5 and out[0, :320].abs().max() == 0
