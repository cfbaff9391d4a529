"""oracle/models.py against tests/golden/wiring_golden.safetensors — outputs of the REFERENCE's own model files
(controlnet_sdv.py, controlnet_sdv_cam_infer.py, controlnet_sdv_bbox.py, unet_spatio_temporal_condition_controlnet.py and
the four block forwards of modified_svd.py), executed unmodified with `diffusers` shimmed by the oracle's leaf blocks
(tests/golden/gen_wiring_golden.py).  Pins the wiring facts the kernels rely on — constructor channel plans and zero-conv
order, the in-loop residual accumulation (multipliers [4,4,4,4,3,3,3,2,2,2,1,1]), the temporal-context interleave
(SURVEY fact 11), the camera / bbox branches and conditioning_scale — against reference CODE rather than a reading of it.

Tolerance: 1e-5 absolute on O(1) values.  The two programs run the same leaf arithmetic in the same order; what differs
is view/reshape/contiguous plumbing, i.e. fp32 round-off of a few ulp (measured 3e-6 when the golden was minted).
"""
import math
from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

from parity_util import make_small_bbox_maps, make_small_inputs, oracle_pair, small_cfg

GOLDEN = Path(__file__).parent / "golden" / "wiring_golden.safetensors"
ATOL = 1e-5


def _summ(t, n=512):
    flat = t.detach().double().reshape(-1)
    stride = max(1, flat.numel() // n)
    return torch.tensor([float(flat.sum()), float(flat.norm())], dtype=torch.float64), flat[::stride][:n].float()


@pytest.mark.parametrize("variant", ["plain", "cam", "bbox"])
def test_oracle_reproduces_reference_wiring(variant):
    g = load_file(str(GOLDEN))
    torch.set_num_threads(4)
    cfg = small_cfg()
    flags = dict(cam=variant == "cam", bbox=variant == "bbox")
    unet, cnet = oracle_pair(cfg, seed=0, **flags)
    inp = make_small_inputs(cfg)
    sigma = 10.0
    x = torch.cat([torch.cat([inp["latents"]] * 2) / (sigma ** 2 + 1) ** 0.5, inp["image_latents"]], dim=2)
    t = torch.tensor(0.25 * math.log(sigma))
    assert torch.equal(x, g["sample"]) and torch.equal(t.reshape(1), g["timestep"])   # same seeded inputs as the generator
    kw = {}
    if flags["cam"]:
        kw["camera_cond"] = inp["camera_cond"]
    if flags["bbox"]:
        kw["controlnet_bbox"] = make_small_bbox_maps(cfg, inp)
    with torch.no_grad():
        down, mid = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"],
                         conditioning_scale=0.8, **kw)
        pred = unet(x, t, inp["image_embeddings"], down_block_additional_residuals=down,
                    mid_block_additional_residual=mid, added_time_ids=inp["added_time_ids"])
    assert len(down) == 12
    for i, r in enumerate(list(down) + [mid]):
        stats, samples = _summ(r)
        want_stats, want_samples = g[f"{variant}.res{i}.stats"], g[f"{variant}.res{i}.samples"]
        assert samples.shape == want_samples.shape
        assert float((samples - want_samples).abs().max()) <= ATOL, (variant, i)
        assert abs(float(stats[1] - want_stats[1])) <= 1e-5 * float(want_stats[1]) + 1e-9, (variant, i)
    want = g[f"{variant}.noise_pred"]
    assert pred.shape == want.shape
    assert float((pred - want).abs().max()) <= ATOL
    assert float(want.abs().max()) > 0.1      # the comparison is not between two zero tensors


def test_variants_are_distinct_in_the_golden():
    g = load_file(str(GOLDEN))
    a, b, c = (g[f"{v}.res0.samples"] for v in ("plain", "cam", "bbox"))
    assert float((a - b).abs().max()) > 1e-4 and float((a - c).abs().max()) > 1e-4
