"""Parity of the image-conditioning branch (posetraj_b200.clip) against the oracle:
  * anti-aliased resize: fp32 kernel against tests/golden/resize_golden.pt (outputs of the reference's own functions)
    and against the oracle on further shapes — absolute error <= 1e-4 (north_star: fp32 elementwise kernels);
  * CLIP vision tower: `image_embeds` relative L2 <= 1e-2 (bf16) against the fp32 oracle on identical weights;
  * `_encode_image` through the pipeline mirror.
"""
import json
import os

import pytest
import torch

from parity_util import rel_l2

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "resize_golden.pt")
SMALL = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, image_size=56,
             patch_size=14, projection_dim=64)


def _record(name, value):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_errors.jsonl", "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


def test_resize_matches_reference_golden(cuda_dev):
    from posetraj_b200.clip import resize_with_antialiasing
    cases = torch.load(GOLDEN)
    g = torch.Generator().manual_seed(0)
    for name, c in cases.items():
        h, w = c["shape"]
        x = torch.rand(1, 3, h, w, generator=g)
        y = resize_with_antialiasing(x.to(cuda_dev), (224, 224)).cpu()
        err = (y[:, :, ::7, ::5] - c["sample"]).abs().max().item()
        _record(f"resize_golden[{name}]", err)
        assert err <= 1e-4, (name, err)
        assert abs(float(y.double().sum()) - c["sum"]) < 0.5, name


@pytest.mark.parametrize("shape,size", [((3, 64, 48), 32), ((3, 31, 200), 56), ((2, 3, 120, 90), 224), ((3, 16, 16), 56)])
def test_resize_matches_oracle(cuda_dev, shape, size):
    from oracle.clip import resize_with_antialiasing as oracle_resize
    from posetraj_b200.clip import resize_with_antialiasing
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(4))
    ref = oracle_resize(x, (size, size))
    out = resize_with_antialiasing(x.to(cuda_dev), (size, size)).cpu()
    err = (out - ref).abs().max().item()
    _record(f"resize[{shape}->{size}]", err)
    assert out.shape == ref.shape and err <= 1e-4, err


def _pair(cuda_dev, act="gelu", **over):
    from oracle.clip import CLIPVisionModelWithProjection as Oracle
    from posetraj_b200.clip import CLIPVisionConfig, CLIPVisionModelWithProjection
    kw = dict(SMALL, hidden_act=act)
    kw.update(over)
    torch.manual_seed(0)
    o = Oracle(**kw).eval()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for name, p in o.named_parameters():
            if "norm" in name:
                p.copy_((1.0 + 0.2 * torch.randn(p.shape, generator=g)) if name.endswith("weight")
                        else 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            if p.dim() > 1:
                p.copy_(p.to(torch.bfloat16).float())
    m = CLIPVisionModelWithProjection(CLIPVisionConfig(**kw), o.state_dict(), cuda_dev)
    return o, m


@pytest.mark.parametrize("act,heads", [("gelu", 4), ("quick_gelu", 2), ("gelu", 1)])
def test_vision_tower_parity(cuda_dev, act, heads):
    o, m = _pair(cuda_dev, act, num_attention_heads=heads)
    x = torch.rand(2, 3, 56, 56, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = o(x)
    out = m(x.to(cuda_dev)).image_embeds
    assert out.shape == ref.shape and out.dtype == torch.float32
    err = rel_l2(out, ref)
    _record(f"clip_tower[{act},{heads}]", err)
    assert err <= 1e-2, err


def test_head_dim_80_and_encode_image(cuda_dev):
    """head_dim 80 (ViT-H's) and the fused resize -> patch rows path, from a 320x576-shaped image."""
    from oracle.clip import encode_image
    o, m = _pair(cuda_dev, hidden_size=320, num_attention_heads=4, intermediate_size=640)
    img = torch.rand(1, 3, 80, 144, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = encode_image(o, img)
    out = m.encode_image(img.to(cuda_dev))
    assert out.shape == ref.shape == (1, 1, 64)
    err = rel_l2(out, ref)
    _record("clip_encode_image[hd80]", err)
    assert err <= 1e-2, err


def test_pipeline_encode_image(cuda_dev):
    import numpy as np
    import PIL.Image
    from oracle.clip import encode_image
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    o, m = _pair(cuda_dev)
    pipe = StableVideoDiffusionPipelineControlNet(image_encoder=m)
    arr = np.random.default_rng(0).integers(0, 256, size=(72, 100, 3), dtype=np.uint8)
    emb = pipe._encode_image(PIL.Image.fromarray(arr), cuda_dev, 1, True)
    assert emb.shape == (2, 1, 64) and float(emb[0].abs().max()) == 0.0
    x = torch.from_numpy(arr.astype(np.float32) / 255.0).permute(2, 0, 1)[None]
    with torch.no_grad():
        ref = encode_image(o, x)
    err = rel_l2(emb[1:], ref)
    _record("pipeline_encode_image", err)
    assert err <= 1e-2, err


def test_argument_errors(cuda_dev):
    _, m = _pair(cuda_dev)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 48, 56, device=cuda_dev))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 56, 56))
