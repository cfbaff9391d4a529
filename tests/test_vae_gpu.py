"""Parity of the VAE mirror (posetraj_b200.vae) against the fp32 CPU oracle on identical weights
and inputs, through the diffusers-shaped API (`encode(...).latent_dist.mode()`, `decode(..., num_frames).sample`).

Tolerance: relative L2 <= 2e-2 on decoded frames / encoded latents (bf16 activations through ~60 convolutions; the
north_star tolerance for one denoise step, a network of similar depth, is 1e-2; the VAE ends in 64-channel layers whose
GroupNorm groups hold only 2 channels, which amplifies bf16 rounding).
"""
import json
import os

import pytest
import torch

from parity_util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-2
SMALL_CH = (64, 64, 128, 128)


def _record(name, value):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_errors.jsonl", "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


def oracle_vae(seed=0, ch=SMALL_CH):
    from oracle.vae import build_vae
    vae = build_vae(seed, block_out_channels=ch)
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        for name, p in vae.named_parameters():
            if "norm" in name.split(".")[-2]:      # non-trivial GroupNorm affine parameters
                p.copy_((1.0 + 0.3 * torch.randn(p.shape, generator=g)) if name.endswith("weight")
                        else 0.1 * torch.randn(p.shape, generator=g))
            if p.dim() > 1:                        # matrices / kernels live in bf16 on the device
                p.copy_(p.to(torch.bfloat16).float())
    return vae


@pytest.fixture(scope="module")
def vae_pair(cuda_dev):
    from posetraj_b200.vae import AutoencoderKLTemporalDecoder, VaeConfig
    o = oracle_vae()
    v = AutoencoderKLTemporalDecoder(VaeConfig(block_out_channels=SMALL_CH), o.state_dict(), cuda_dev)
    return o, v


@pytest.mark.parametrize("batch,frames,h,w", [(1, 3, 8, 12), (2, 2, 8, 8), (1, 1, 4, 8)])
def test_decode_parity(vae_pair, cuda_dev, batch, frames, h, w):
    o, v = vae_pair
    g = torch.Generator().manual_seed(11)
    z = torch.randn(batch * frames, 4, h, w, generator=g) / 0.18215
    with torch.no_grad():
        ref = o.decode(z, frames)
    out = v.decode(z.to(cuda_dev), num_frames=frames).sample
    assert out.shape == ref.shape and out.dtype == torch.float32
    err = rel_l2(out, ref)
    _record(f"vae_decode[{batch}x{frames}x{h}x{w}]", err)
    assert err <= TOL, err


def test_decode_odd_token_count(vae_pair, cuda_dev):
    """h*w = 40 is not a multiple of 64: the attention's padded key columns must stay out of the softmax."""
    o, v = vae_pair
    z = torch.randn(2, 4, 5, 8, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = o.decode(z, 2)
    out = v.decode(z.to(cuda_dev), num_frames=2).sample
    err = rel_l2(out, ref)
    _record("vae_decode[odd tokens]", err)
    assert err <= TOL, err


@pytest.mark.parametrize("n,H,W", [(1, 64, 96), (2, 32, 32)])
def test_encode_parity(vae_pair, cuda_dev, n, H, W):
    o, v = vae_pair
    x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(3)) * 2 - 1
    with torch.no_grad():
        ref = o.encode_mode(x)
    dist = v.encode(x.to(cuda_dev)).latent_dist
    out = dist.mode()
    assert out.shape == ref.shape
    err = rel_l2(out, ref)
    _record(f"vae_encode[{n}x{H}x{W}]", err)
    assert err <= TOL, err
    assert dist.sample(generator=torch.Generator(device=cuda_dev).manual_seed(0)).shape == ref.shape


def test_chunked_decode_and_from_pretrained(vae_pair, cuda_dev, tmp_path):
    """`decode_latents` decodes `decode_chunk_size` frames at a time (pipeline...controlnet.py:238-246): the temporal
    layers only see one chunk; and a checkpoint written in the diffusers directory layout loads back."""
    from posetraj_b200.checkpoint import save_pretrained
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.vae import AutoencoderKLTemporalDecoder
    o, v = vae_pair
    save_pretrained(v.state_dict(), v.cfg, str(tmp_path / "vae"), "AutoencoderKLTemporalDecoder")
    v2 = AutoencoderKLTemporalDecoder.from_pretrained(str(tmp_path), subfolder="vae", device=cuda_dev)
    assert v2.cfg == v.cfg and v2.config.scaling_factor == 0.18215
    pipe = StableVideoDiffusionPipelineControlNet(vae=v2)
    lat = torch.randn(1, 5, 4, 8, 8, generator=torch.Generator().manual_seed(21))
    out = pipe.decode_latents(lat.to(cuda_dev), 5, decode_chunk_size=2)          # chunks of 2, 2, 1 frames
    assert out.shape == (1, 3, 5, 64, 64) and out.dtype == torch.float32
    z = lat[0] / 0.18215
    with torch.no_grad():
        ref = torch.cat([o.decode(z[0:2], 2), o.decode(z[2:4], 2), o.decode(z[4:5], 1)])
    err = rel_l2(out[0].permute(1, 0, 2, 3), ref)
    _record("vae_chunked_decode", err)
    assert err <= TOL, err


def test_argument_errors(vae_pair, cuda_dev):
    _, v = vae_pair
    with pytest.raises(ValueError):
        v.decode(torch.zeros(3, 4, 8, 8, device=cuda_dev), num_frames=2)
    with pytest.raises(ValueError):
        v.decode(torch.zeros(2, 3, 8, 8, device=cuda_dev), num_frames=2)
    with pytest.raises(RuntimeError):
        v.decode(torch.zeros(2, 4, 8, 8), num_frames=2)
    with pytest.raises(ValueError):
        v.encode(torch.zeros(1, 3, 36, 32, device=cuda_dev))


def test_pipeline_decodes_frames(cuda_dev):
    """`output_type="pt"` through the pipeline mirror: the denoised latents go through decode_latents + tensor2vid
    (pipeline...controlnet.py:585-592) and equal the oracle VAE applied to the pipeline's own latents."""
    from parity_util import make_small_inputs, oracle_pair, small_cfg
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.vae import AutoencoderKLTemporalDecoder, VaeConfig
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=0)
    unet = UNetSpatioTemporalConditionControlNetModel(cfg, o_unet.state_dict(), cuda_dev)
    cnet = ControlNetSDVModel(cfg, o_cnet.state_dict(), cuda_dev)
    o_vae = oracle_vae(seed=2)
    vae = AutoencoderKLTemporalDecoder(VaeConfig(block_out_channels=SMALL_CH), o_vae.state_dict(), cuda_dev)
    pipe = StableVideoDiffusionPipelineControlNet(vae=vae, unet=unet, controlnet=cnet)
    inp = make_small_inputs(cfg, h=16, w=16)
    kw = dict(controlnet_condition=inp["controlnet_condition"][0], height=128, width=128, num_frames=cfg.num_frames,
              num_inference_steps=2, image_embeddings=inp["image_embeddings"])
    lat = pipe(latents=inp["latents"] / 700.0, image_latents=inp["image_latents"][:, 0], output_type="latent", **kw).frames
    out = pipe(latents=inp["latents"] / 700.0, image_latents=inp["image_latents"][:, 0], output_type="pt", **kw).frames
    assert isinstance(out, list) and len(out) == 1 and out[0].shape == (cfg.num_frames, 3, 128, 128)
    with torch.no_grad():
        ref = o_vae.decode(lat[0].float().cpu() / 0.18215, cfg.num_frames)
    ref = (ref / 2 + 0.5).clamp(0, 1)
    err = rel_l2(out[0], ref)
    _record("pipeline_pt_frames", err)
    assert err <= TOL, err
    # conditioning image through the VAE encoder instead of image_latents=
    img = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(9))
    out2 = pipe(image=img, latents=inp["latents"] / 700.0, output_type="np", generator=torch.Generator().manual_seed(1), **kw).frames
    assert out2[0].shape == (cfg.num_frames, 128, 128, 3)
