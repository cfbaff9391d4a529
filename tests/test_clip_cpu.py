"""CPU checks of the image-conditioning row: the resize oracle against golden vectors produced
by the reference's own functions, the CLIP oracle against the installed `transformers` implementation, key tree /
parameter count of the mirror, and that the product refuses to run without CUDA."""
import math
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "resize_golden.pt")


def test_resize_oracle_matches_reference_golden():
    """tests/golden/resize_golden.pt was produced by executing /root/reference/pipeline/...controlnet.py:602-712."""
    from oracle.clip import resize_with_antialiasing
    cases = torch.load(GOLDEN)
    g = torch.Generator().manual_seed(0)
    for name, c in cases.items():
        h, w = c["shape"]
        x = torch.rand(1, 3, h, w, generator=g)
        y = resize_with_antialiasing(x)
        assert y.shape == (1, 3, 224, 224)
        assert torch.allclose(y[:, :, ::7, ::5], c["sample"], atol=1e-6, rtol=0), name
        assert abs(float(y.double().sum()) - c["sum"]) < 1e-3, name


def test_clip_oracle_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from oracle.clip import CLIPVisionModelWithProjection as Oracle
    for act in ("gelu", "quick_gelu"):
        kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, image_size=56,
                  patch_size=14, projection_dim=64, hidden_act=act)
        torch.manual_seed(0)
        hf = transformers.CLIPVisionModelWithProjection(transformers.CLIPVisionConfig(**kw)).eval()
        ours = Oracle(**kw).eval()
        sd = {k: v for k, v in hf.state_dict().items() if "position_ids" not in k}
        ours.load_state_dict(sd)   # identical key tree
        x = torch.rand(2, 3, 56, 56, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            ref = hf(pixel_values=x).image_embeds
            out = ours(x)
        assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), (act, (out - ref).abs().max())


def test_key_tree_and_parameter_count():
    from oracle.clip import CLIPVisionModelWithProjection as Oracle
    from posetraj_b200.clip import CLIPVisionConfig, clip_param_shapes
    small = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, image_size=56)
    shapes = clip_param_shapes(CLIPVisionConfig(**small))
    sd = Oracle(**small).state_dict()
    assert set(sd) == set(shapes)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k
    # ViT-H/14 with projection (the SVD image_encoder): 632.08 M parameters
    full = clip_param_shapes(CLIPVisionConfig())
    assert sum(math.prod(s) for s in full.values()) == 632_076_800


def test_clip_refuses_cpu():
    from posetraj_b200.clip import CLIPVisionConfig, CLIPVisionModelWithProjection, resize_with_antialiasing
    with pytest.raises(RuntimeError):
        CLIPVisionModelWithProjection(CLIPVisionConfig(), {}, device="cpu")
    with pytest.raises(RuntimeError):
        resize_with_antialiasing(torch.zeros(1, 3, 32, 32))


def test_load_hf_checkpoint_directory(tmp_path):
    """`from_pretrained` reads the Hugging Face layout of SVD's image_encoder/ (config.json + model.safetensors, or
    pytorch_model.bin), ignoring keys the mirror does not use (e.g. `position_ids` buffers, `_name_or_path`)."""
    import json
    from safetensors.torch import save_file
    from oracle.clip import CLIPVisionModelWithProjection as Oracle
    from posetraj_b200.clip import clip_param_shapes, load_clip_checkpoint
    kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, image_size=56,
              patch_size=14, projection_dim=64, hidden_act="quick_gelu")
    sd = {k: v.detach().clone() for k, v in Oracle(**kw).state_dict().items()}
    sd["vision_model.embeddings.position_ids"] = torch.arange(17)[None]
    d = tmp_path / "image_encoder"
    d.mkdir()
    json.dump(dict(kw, _name_or_path="x", architectures=["CLIPVisionModelWithProjection"], dropout=0.0), open(d / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "model.fp16.safetensors"))
    cfg, loaded = load_clip_checkpoint(str(tmp_path), subfolder="image_encoder", variant="fp16")
    assert cfg.hidden_act == "quick_gelu" and cfg.hidden_size == 128 and cfg.projection_dim == 64
    shapes = clip_param_shapes(cfg)
    assert set(shapes) <= set(loaded) and all(tuple(loaded[k].shape) == tuple(s) for k, s in shapes.items())
    with pytest.raises(FileNotFoundError):
        load_clip_checkpoint(str(tmp_path), subfolder="image_encoder")        # no un-suffixed weight file
    torch.save(sd, str(d / "pytorch_model.bin"))
    cfg2, loaded2 = load_clip_checkpoint(str(tmp_path), subfolder="image_encoder", num_hidden_layers=2)
    assert torch.equal(loaded2["visual_projection.weight"], sd["visual_projection.weight"])
