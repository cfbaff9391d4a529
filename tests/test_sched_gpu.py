"""Fused CFG + Euler kernel vs. the literal reference formulas in torch fp32 (tolerance 1e-4, north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def karras_sigmas(n=25, smin=0.002, smax=700.0, rho=7.0):
    ramp = torch.linspace(0, 1, n, dtype=torch.float64)
    s = (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho
    return torch.cat([s.float(), torch.zeros(1)])


@pytest.mark.parametrize("step", [0, 12, 24])
@pytest.mark.parametrize("nchw", [False, True])
def test_cfg_euler(cuda_dev, step, nchw):
    from posetraj_b200.ops import CfgEuler
    torch.manual_seed(step)
    Fr, Cc, H, W = 14, 4, 40, 72
    sig = karras_sigmas().cuda()
    lat = (torch.randn(Fr, Cc, H, W, device="cuda") * (sig[step] ** 2 + 1).sqrt()).contiguous()
    lat0 = lat.clone()
    img = torch.randn(2, Fr, Cc, H, W, device="cuda")
    img[0] = 0
    g = torch.linspace(1, 3, Fr, device="cuda")
    si = torch.tensor([step], device="cuda", dtype=torch.int32)
    pred = torch.randn(2, Fr, Cc, H, W, device="cuda")
    if nchw:
        pred_arg = pred.contiguous()
    else:
        pred = pred.to(torch.bfloat16).float()
        pred_arg = pred.permute(0, 1, 3, 4, 2).reshape(2 * Fr * H * W, Cc).contiguous().to(torch.bfloat16)
    nxt = torch.zeros(2 * Fr * (H + 1) * (W + 1), 64, device="cuda", dtype=torch.bfloat16)
    CfgEuler(noise_pred=pred_arg, latents=lat, guidance=g, sigmas=sig, step_index=si, next_in=nxt,
             image_latents=img, pred_nchw_f32=nchw).launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # reference: pipeline :567-569 + scheduler.step :481-517 (fp32)
    s, sn = sig[step], sig[step + 1]
    v = pred[0] + g.view(Fr, 1, 1, 1) * (pred[1] - pred[0])
    x0 = v * (-s / (s ** 2 + 1) ** 0.5) + lat0 / (s ** 2 + 1)
    ref = lat0 + (lat0 - x0) / s * (sn - s)
    err = ((lat - ref).norm() / ref.norm()).item()
    assert err < 1e-4, err
    # next model input: cat(x / sqrt(sn^2+1), image_latents) in the zero-haloed NHWC layout
    nin = nxt.view(2, Fr, H + 1, W + 1, 64).float()
    want = torch.cat([(ref / (sn ** 2 + 1) ** 0.5).expand(2, Fr, Cc, H, W), img], 2).permute(0, 1, 3, 4, 2)
    got = nin[:, :, :H, :W, :2 * Cc]
    assert ((got - want).norm() / want.norm()).item() < 4e-3
    assert nin[:, :, H].abs().max() == 0 and nin[:, :, :, W].abs().max() == 0 and nin[..., 2 * Cc:].abs().max() == 0
