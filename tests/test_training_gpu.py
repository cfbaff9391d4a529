"""Backward / loss / optimizer kernels of the configs[3] training step against torch autograd on the
same GPU in fp32 — operator by operator, then two assembled blocks (SpatioTemporalResBlock, LayerNorm + GEGLU feed-forward)
against autograd of the oracle's modules.

Tolerances: inputs are bf16-rounded on both sides; gradients leave the kernels as bf16 (activations) or fp32 (parameters).
Operator level: rel-L2 <= 5e-3.  Block level (4-8 chained bf16 hand-offs): <= 2e-2 — the bound VERDICT r1 set for the
first training milestone.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL_OP = 5e-3
TOL_BLOCK = 2e-2


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def test_linear_dgrad_and_wgrad(cuda_dev):
    from posetraj_b200 import training as T
    torch.manual_seed(0)
    for (M, N, K) in [(1000, 128, 64), (5000, 320, 320), (20000, 640, 1280)]:
        a, w, d = rnd(M, K), rnd(N, K, scale=1 / math.sqrt(K)), rnd(M, N)
        da = T.linear_dgrad(d, w)
        dw = T.wgrad(d, a)
        db = T.colsum(d)[0]
        torch.cuda.synchronize()
        assert rel_l2(da, d.float() @ w.float()) < TOL_OP
        assert rel_l2(dw, d.float().t() @ a.float()) < TOL_OP
        assert rel_l2(db, d.float().sum(0)) < TOL_OP
    # a fixed split count must give the same bits twice (deterministic fold)
    a, d = rnd(7000, 192), rnd(7000, 256)
    x, y = T.wgrad(d, a, splits=5), T.wgrad(d, a, splits=5)
    torch.cuda.synchronize()
    assert torch.equal(x, y)


@pytest.mark.parametrize("n,H,W,Cin,Cout", [(3, 16, 24, 64, 128), (2, 10, 18, 320, 320)])
def test_conv3x3_backward(cuda_dev, n, H, W, Cin, Cout):
    """dgrad / wgrad of the implicit-GEMM 3x3 conv in the zero-haloed row space against autograd of F.conv2d."""
    from posetraj_b200 import ops, training as T
    torch.manual_seed(1)
    x = rnd(n, Cin, H, W)
    wt = rnd(Cout, Cin, 3, 3, scale=1 / math.sqrt(9 * Cin))
    dy = rnd(n, Cout, H, W)
    xr, wr = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(dy.float())
    # library layouts
    tok = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
    x_h, dy_h = T.to_halo(tok(x), n, H, W), T.to_halo(tok(dy), n, H, W)
    w_k = wt.float().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(torch.bfloat16).contiguous()
    taps = ops.conv3x3_taps(W)
    dx_h = T.linear_dgrad(dy_h, w_k, taps=taps)
    dw = T.wgrad(dy_h, x_h, taps=taps)
    torch.cuda.synchronize()
    dx = dx_h.view(n, H + 1, W + 1, Cin)[:, :H, :W].permute(0, 3, 1, 2)
    assert rel_l2(dx, xr.grad) < TOL_OP
    assert rel_l2(dw.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2), wr.grad) < TOL_OP


def test_temporal_conv_backward(cuda_dev):
    from posetraj_b200 import training as T
    torch.manual_seed(2)
    B, Fr, HW, Cc = 2, 5, 96, 128
    x = rnd(B, Cc, Fr, HW, 1)
    wt = rnd(Cc, Cc, 3, 1, 1, scale=1 / math.sqrt(3 * Cc))
    dy = rnd(B, Cc, Fr, HW, 1)
    xr, wr = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    F.conv3d(xr, wr, padding=(1, 0, 0)).backward(dy.float())
    tok = lambda t: t.permute(0, 2, 3, 4, 1).reshape(-1, Cc).contiguous()       # rows (b, f, s)
    w_k = wt.float().reshape(Cc, Cc, 3).permute(0, 2, 1).reshape(Cc, 3 * Cc).to(torch.bfloat16).contiguous()
    taps = (-HW, 0, HW)
    dx = T.linear_dgrad(tok(dy), w_k, taps=taps, batches=B)
    dw = T.wgrad(tok(dy), tok(x), taps=taps, batches=B)
    torch.cuda.synchronize()
    assert rel_l2(dx, tok(xr.grad)) < TOL_OP
    assert rel_l2(dw.view(Cc, 3, Cc).permute(0, 2, 1), wr.grad.reshape(Cc, Cc, 3)) < TOL_OP


@pytest.mark.parametrize("stat5d,cat,halo", [(False, False, True), (True, False, False), (False, True, True)])
def test_groupnorm_backward(cuda_dev, stat5d, cat, halo):
    from posetraj_b200 import training as T
    torch.manual_seed(3)
    n, H, W, C0, C1 = 4, 8, 12, 64, (64 if cat else 0)
    Cc = C0 + C1
    x = rnd(n * H * W, Cc)
    gamma, beta = torch.randn(Cc, device="cuda") * 0.3 + 1.0, torch.randn(Cc, device="cuda") * 0.2
    dout = rnd(n * H * W, Cc)
    rps = (2 * H * W) if stat5d else H * W
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    S = x.shape[0] // rps
    xs = xr.view(S, rps, 32, Cc // 32)
    mu, var = xs.mean(dim=(1, 3), keepdim=True), xs.var(dim=(1, 3), unbiased=False, keepdim=True)
    y = ((xs - mu) * (var + 1e-6).rsqrt()).view(-1, Cc) * gr + br
    F.silu(y).backward(dout.float())
    d_in = T.to_halo(dout, n, H, W) if halo else dout
    dx0, dx1, dgb = T.groupnorm_backward(x[:, :C0].contiguous(), d_in, gamma, beta, rows_per_stat=rps, eps=1e-6,
                                         x1=x[:, C0:].contiguous() if cat else None, halo=(H, W) if halo else None)
    torch.cuda.synchronize()
    dx = torch.cat([dx0, dx1], 1) if cat else dx0
    assert rel_l2(dx, xr.grad) < TOL_OP
    assert rel_l2(dgb[0], gr.grad) < TOL_OP and rel_l2(dgb[1], br.grad) < TOL_OP


def test_layernorm_geglu_backward(cuda_dev):
    from posetraj_b200 import training as T
    torch.manual_seed(4)
    rows, Cc = 3000, 320
    x, dout = rnd(rows, Cc), rnd(rows, Cc)
    gamma, beta = torch.randn(Cc, device="cuda") * 0.3 + 1.0, torch.randn(Cc, device="cuda") * 0.2
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (Cc,), gr, br, 1e-5).backward(dout.float())
    dx, dgb = T.layernorm_backward(x, dout, gamma)
    torch.cuda.synchronize()
    assert rel_l2(dx, xr.grad) < TOL_OP and rel_l2(dgb[0], gr.grad) < TOL_OP and rel_l2(dgb[1], br.grad) < TOL_OP
    h, d2 = rnd(rows, 2 * Cc), rnd(rows, Cc)
    hr = h.float().requires_grad_(True)
    v, g = hr.chunk(2, dim=-1)
    out_ref = v * F.gelu(g)
    out_ref.backward(d2.float())
    out = T.geglu_forward(h)
    dh = T.geglu_backward(h, d2)
    torch.cuda.synchronize()
    assert rel_l2(out, out_ref) < TOL_OP and rel_l2(dh, hr.grad) < TOL_OP


def test_edm_loss_and_gradient(cuda_dev):
    """pt_edm_loss against the reference's loss lines (train...cam_concat.py:1417-1436) under autograd, main pass and the
    F = 1 'spatial' pass accumulated with weight 0.5 (:1438-1462)."""
    from posetraj_b200 import training as T
    torch.manual_seed(5)
    B, Fr, Cc, H, W = 2, 3, 4, 16, 24
    lat, noise = torch.randn(B, Fr, Cc, H, W, device="cuda"), torch.randn(B, Fr, Cc, H, W, device="cuda")
    sig = torch.tensor([1.7, 0.3], device="cuda")
    s = sig.view(B, 1, 1, 1, 1)
    noisy = lat + noise * s
    pred = rnd(B * Fr * H * W, Cc)
    pred_sp = rnd(B * 1 * H * W, Cc)
    pr = pred.float().view(B, Fr, H, W, Cc).permute(0, 1, 4, 2, 3).requires_grad_(True)
    ps = pred_sp.float().view(B, 1, H, W, Cc).permute(0, 1, 4, 2, 3).requires_grad_(True)
    c_out, c_skip, wgt = -s / (s ** 2 + 1) ** 0.5, 1 / (s ** 2 + 1), (1 + s ** 2) * s ** -2.0
    main = torch.mean((wgt * (pr * c_out + c_skip * noisy - lat) ** 2).reshape(B, -1), dim=1).mean()
    ran = 1
    spat = torch.mean((wgt[:, 0] * (ps[:, 0] * c_out[:, 0] + c_skip[:, 0] * noisy[:, ran] - lat[:, ran]) ** 2).reshape(B, -1), dim=1).mean()
    (main + 0.5 * spat).backward()
    loss, dpred = T.edm_loss(pred, noisy, lat, sig)
    loss, dsp = T.edm_loss(pred_sp, noisy, lat, sig, weight=0.5, frame=ran, loss=loss)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(main + 0.5 * spat)) < 1e-4 * float(main + 0.5 * spat)
    tok = lambda t: t.permute(0, 1, 3, 4, 2).reshape(-1, Cc)
    assert rel_l2(dpred, tok(pr.grad)) < TOL_OP and rel_l2(dsp, tok(ps.grad)) < TOL_OP


def test_adamw_matches_torch(cuda_dev):
    from posetraj_b200 import training as T
    torch.manual_seed(6)
    sizes = [1000, 77, 4096]
    params = [torch.randn(n, device="cuda") for n in sizes]
    ref = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    gb = T.GradientBuckets(sizes, "cuda", bucket_mb=0.01)
    mine = T.AdamW(gb, params, work=params, lr=1e-3)
    for step in range(3):
        for i, r in enumerate(ref):
            g = torch.randn_like(r)
            r.grad = g.clone()
            gb.view(i).copy_(g)
            gb.ready(i)
        gb.finish()
        opt.step()
        mine.step()
    torch.cuda.synchronize()
    for i, r in enumerate(ref):
        assert rel_l2(mine.param(i), r.detach()) < 1e-5


def test_feedforward_block_forward_backward(cuda_dev):
    """LayerNorm -> GEGLU -> Linear + residual (BasicTransformerBlock norm3 + ff) against autograd of the oracle modules."""
    from oracle.svd_blocks import FeedForward
    from posetraj_b200 import training as T
    torch.manual_seed(7)
    rows, Cc = 2880, 320
    ff = FeedForward(Cc).cuda()
    ln = torch.nn.LayerNorm(Cc).cuda()
    with torch.no_grad():
        ln.weight.add_(torch.randn(Cc, device="cuda") * 0.1)
        ln.bias.add_(torch.randn(Cc, device="cuda") * 0.1)
        for p in ff.parameters():
            if p.dim() > 1:
                p.copy_(p.to(torch.bfloat16).float())
    x, dout = rnd(rows, Cc), rnd(rows, Cc)
    xr = x.float().requires_grad_(True)
    ref = xr + ff(ln(xr))
    ref.backward(dout.float())
    tr = T.FeedForwardTrainer(ln.weight.detach(), ln.bias.detach(), ff.net[0].proj.weight.detach().to(torch.bfloat16),
                              ff.net[0].proj.bias.detach(), ff.net[2].weight.detach().to(torch.bfloat16), ff.net[2].bias.detach())
    out = tr.forward(x)
    dx, g = tr.backward(dout)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < TOL_BLOCK
    assert rel_l2(dx, xr.grad) < TOL_BLOCK
    want = {"w1": ff.net[0].proj.weight.grad, "b1": ff.net[0].proj.bias.grad, "w2": ff.net[2].weight.grad, "b2": ff.net[2].bias.grad,
            "ln_w": ln.weight.grad, "ln_b": ln.bias.grad}
    errs = {k: rel_l2(g[k], v) for k, v in want.items()}
    assert max(errs.values()) < TOL_BLOCK, errs


@pytest.mark.parametrize("cin,cout", [(128, 128), (64, 128)])
def test_resblock_forward_backward(cuda_dev, cin, cout):
    """SpatioTemporalResBlock forward + backward on the library against autograd of oracle.svd_blocks (fp32): every
    parameter gradient of the block, the input gradient and the time-embedding gradients."""
    from oracle.svd_blocks import SpatioTemporalResBlock
    from posetraj_b200 import training as T
    torch.manual_seed(8)
    B, Fr, H, W, temb = 2, 3, 8, 12, 256
    blk = SpatioTemporalResBlock(cin, cout, temb, 1e-6).cuda()
    with torch.no_grad():
        for name, p in blk.named_parameters():
            if p.dim() > 1:
                p.copy_(p.to(torch.bfloat16).float())
            elif "norm" in name:
                p.add_(torch.randn_like(p) * 0.1)
        blk.time_mixer.mix_factor.fill_(0.3)
    x = rnd(B * Fr, cin, H, W)
    emb = torch.randn(B, temb, device="cuda")
    dout = rnd(B * Fr, cout, H, W)
    xr = x.float().requires_grad_(True)
    er = emb.clone().requires_grad_(True)
    ioi = torch.zeros(B, Fr, device="cuda")
    ref = blk(xr, er.repeat_interleave(Fr, dim=0), ioi)
    ref.backward(dout.float())
    # the library takes the two time-embedding projections as inputs (engine.py batches them into one GEMV per step)
    sb, tb = blk.spatial_res_block, blk.temporal_res_block
    with torch.no_grad():
        temb_s = sb.time_emb_proj(F.silu(emb)).contiguous()
        temb_t = tb.time_emb_proj(F.silu(emb)).contiguous()
    params = {k: v.detach() for k, v in blk.named_parameters()}
    tr = T.ResBlockTrainer(params, B=B, F=Fr, H=H, W=W, eps=1e-6)
    tok = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
    out = tr.forward(tok(x), temb_s, temb_t)
    dx, g = tr.backward(tok(dout))
    torch.cuda.synchronize()
    assert rel_l2(out, tok(ref)) < TOL_BLOCK
    assert rel_l2(dx, tok(xr.grad)) < TOL_BLOCK, rel_l2(dx, tok(xr.grad))
    errs = {}
    for name, p in blk.named_parameters():
        if "time_emb_proj" in name:
            continue
        got = g[name]
        if name.endswith("conv1.weight") or name.endswith("conv2.weight"):
            if name.startswith("spatial"):
                got = got.view(p.shape[0], 3, 3, -1).permute(0, 3, 1, 2)
            else:
                got = got.view(p.shape[0], 3, -1).permute(0, 2, 1).reshape(p.shape)
        errs[name] = rel_l2(got.reshape(p.shape), p.grad)
    # time-embedding path: d loss / d (time_emb_proj output) per batch row, then the reference's own chain rule
    for which, mod in (("d_temb_s", sb), ("d_temb_t", tb)):
        want_w = mod.time_emb_proj.weight.grad
        got_w = g[which].t() @ F.silu(emb)
        errs[which] = rel_l2(got_w, want_w)
    assert max(errs.values()) < TOL_BLOCK, errs
