"""CPU checks of the VAE row: key tree / parameter count of the mirror against the oracle,
oracle invariants, the image processor, and that the product refuses to run without CUDA."""
import math

import numpy as np
import pytest
import torch


def test_key_tree_and_parameter_count():
    from oracle.vae import build_vae
    from posetraj_b200.vae import VaeConfig, vae_param_shapes
    shapes = vae_param_shapes(VaeConfig())
    sd = build_vae(0).state_dict()
    assert set(sd) == set(shapes)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k
    # the SVD VAE (AutoencoderKLTemporalDecoder, block_out_channels (128, 256, 512, 512)): 97.7 M parameters
    assert sum(math.prod(s) for s in shapes.values()) == 97_742_847


def test_oracle_shapes_and_temporal_coupling():
    from oracle.vae import build_vae
    vae = build_vae(3, block_out_channels=(64, 64, 64, 64))
    g = torch.Generator().manual_seed(0)
    z = torch.randn(3, 4, 8, 8, generator=g)
    with torch.no_grad():
        full = vae.decode(z, 3)
        assert full.shape == (3, 3, 64, 64)
        # frames are coupled through the temporal layers: decoding them one at a time gives a different answer ...
        single = torch.cat([vae.decode(z[i:i + 1], 1) for i in range(3)])
        assert (full - single).abs().max() > 1e-4
        # ... and two videos in one call do not see each other
        z2 = torch.cat([z, torch.randn(3, 4, 8, 8, generator=g)])
        both = vae.decode(z2, 3)
        assert torch.allclose(both[:3], full, atol=1e-5)
        lat = vae.encode_mode(torch.randn(2, 3, 64, 64, generator=g))
        assert lat.shape == (2, 4, 8, 8)


def test_oracle_downsample_is_pad_right_bottom():
    from oracle.vae import DownEncoderBlock2D
    torch.manual_seed(0)
    blk = DownEncoderBlock2D(64, 64, add_downsample=True, num_layers=0)
    x = torch.randn(1, 64, 6, 8)
    with torch.no_grad():
        y = blk(x)
        w, b = blk.downsamplers[0].conv.weight, blk.downsamplers[0].conv.bias
        # out(1, 2) = sum_k w[ky, kx] x(2 + ky, 4 + kx); out(2, 3) touches the padded row 6 / column 8 (zeros)
        ref = (w[:, :, :, :] * x[0, :, 2:5, 4:7][None]).sum((1, 2, 3)) + b
        assert torch.allclose(y[0, :, 1, 2], ref, atol=1e-5)
        patch = torch.zeros(64, 3, 3)
        patch[:, :2, :2] = x[0, :, 4:6, 6:8]
        ref = (w * patch[None]).sum((1, 2, 3)) + b
        assert torch.allclose(y[0, :, 2, 3], ref, atol=1e-5)


def test_image_processor_roundtrip():
    import PIL.Image
    from posetraj_b200.pipeline import VaeImageProcessor, tensor2vid
    proc = VaeImageProcessor()
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, size=(32, 48, 3), dtype=np.uint8)
    x = proc.preprocess(PIL.Image.fromarray(arr), height=32, width=48)
    assert x.shape == (1, 3, 32, 48) and x.min() >= -1 and x.max() <= 1
    back = proc.postprocess(x, "pil")
    assert np.array_equal(np.asarray(back[0]), arr)
    assert proc.postprocess(x, "np").shape == (1, 32, 48, 3)
    vid = tensor2vid(x.permute(1, 0, 2, 3)[None], proc, "pt")   # [B, C, F, H, W]
    assert len(vid) == 1 and vid[0].shape == (1, 3, 32, 48)
    with pytest.raises(ValueError):
        proc.postprocess(x, "gif")
    # tensors already in [-1, 1] are passed through
    t = torch.rand(2, 3, 8, 8) * 2 - 1
    assert torch.equal(proc.preprocess(t), t)


def test_vae_refuses_cpu():
    from posetraj_b200.vae import AutoencoderKLTemporalDecoder, VaeConfig
    with pytest.raises(RuntimeError):
        AutoencoderKLTemporalDecoder(VaeConfig(), {}, device="cpu")


def test_controlnet_condition_from_pil_list():
    """The reference scripts hand the pipeline a list of PIL trajectory maps (run_inference_vipseg_json_repro.py:451);
    `preprocess` (pipeline...controlnet.py:500-503) turns them into [2, F, 3, H, W] in [-1, 1]."""
    import PIL.Image
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    pipe = StableVideoDiffusionPipelineControlNet()
    rng = np.random.default_rng(1)
    imgs = [PIL.Image.fromarray(rng.integers(0, 256, size=(32, 48, 3), dtype=np.uint8)) for _ in range(3)]
    cond = pipe.prepare_controlnet_condition(imgs, 32, 48)
    assert cond.shape == (2, 3, 3, 32, 48) and cond.min() >= -1 and cond.max() <= 1
    assert torch.equal(cond[0], cond[1])
    want = torch.from_numpy(np.asarray(imgs[1], dtype=np.float32) / 255.0).permute(2, 0, 1) * 2 - 1
    assert torch.allclose(cond[0, 1], want)
    t = torch.rand(3, 3, 32, 48) * 2 - 1
    assert torch.equal(pipe.prepare_controlnet_condition(t, 32, 48)[1], t)
    with pytest.raises(ValueError):
        pipe.prepare_controlnet_condition(None, 32, 48)
    with pytest.raises(ValueError):
        pipe.prepare_controlnet_condition(torch.zeros(3, 4, 32, 48), 32, 48)
