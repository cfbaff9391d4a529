"""CPU checks (-m "not gpu"): the oracle against the reference-generated golden vectors and the derived known
answers of SURVEY.md Appendix D, the structural invariants of SURVEY.md §4, and the host-side logic of the product
package (config / key tree / FLOP counter / schedule) that needs no GPU."""
import json
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from parity_util import make_small_inputs, oracle_pair, rel_l2, small_cfg

GOLDEN = Path(__file__).parent / "golden"


# ---------------------------------------------------------------------------------------------------------------
# scheduler: oracle and product host logic vs the reference file's own outputs (tests/golden/gen_scheduler_golden.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sched_golden():
    return json.loads((GOLDEN / "scheduler_golden.json").read_text())


def test_oracle_scheduler_matches_reference_file(sched_golden):
    from oracle.scheduler import EulerKarrasOracle
    for case in sched_golden["cases"]:
        n = case["num_inference_steps"]
        s = EulerKarrasOracle()
        s.set_timesteps(n)
        assert torch.equal(s.sigmas, torch.tensor(case["sigmas"], dtype=torch.float32))
        assert torch.equal(s.timesteps, torch.tensor(case["timesteps"], dtype=torch.float32))
        assert float(s.init_noise_sigma) == case["init_noise_sigma"]
        g = torch.Generator().manual_seed(case["seed"])
        x = torch.randn(case["shape"], generator=g) * s.init_noise_sigma
        scaled = {e["i"]: e["scale_model_input"] for e in case["scaled"]}
        steps = {e["i"]: e["prev_sample"] for e in case["steps"]}
        for i, t in enumerate(s.timesteps):
            xin = s.scale_model_input(x, t)
            v = torch.randn(x.shape, generator=g)
            x = s.step(v, t, x)
            if i in scaled:
                assert torch.equal(xin.flatten(), torch.tensor(scaled[i]))   # bit-exact: same fp32 op order
                assert torch.equal(x.flatten(), torch.tensor(steps[i]))
        assert torch.equal(x.flatten(), torch.tensor(case["final"]))


def test_product_sigma_table_matches_reference_file(sched_golden):
    from posetraj_b200.scheduler import EulerDiscreteScheduler
    for case in sched_golden["cases"]:
        n = case["num_inference_steps"]
        sig = torch.from_numpy(EulerDiscreteScheduler.karras_sigmas(n, 0.002, 700.0)).to(torch.float32)
        assert torch.equal(sig, torch.tensor(case["sigmas"][:-1], dtype=torch.float32))


def test_appendix_d_known_answers():
    """derived fp32 known answers of the 25-step schedule."""
    from oracle.scheduler import EulerKarrasOracle
    s = EulerKarrasOracle()
    s.set_timesteps(25)
    sig = s.sigmas.double().numpy()
    np.testing.assert_allclose(sig[:5], [700, 545.72925, 421.56912, 322.45367, 244.02307], rtol=2e-7)
    np.testing.assert_allclose(sig[20:25], [0.15740465, 0.06639908, 0.02480258, 0.0078825, 0.002], rtol=2e-6)
    assert sig[25] == 0.0
    np.testing.assert_allclose(s.timesteps[:3].numpy(), [1.63777, 1.5755308, 1.510996], rtol=1e-6)
    np.testing.assert_allclose(s.timesteps[-3:].numpy(), [-0.9242019, -1.2107776, -1.553652], rtol=1e-6)
    assert abs(float(s.init_noise_sigma) - 700.000732) < 1e-3
    # Euler closed form x <- a x + b v
    def ab(i):
        sg, sn = sig[i], sig[i + 1]
        c = sg * sg + 1
        a = 1 + (sn - sg) / sg * (1 - 1 / c)
        b = (sn - sg) / sg * (sg / math.sqrt(c))
        return a, b
    np.testing.assert_allclose(ab(0), (0.77961366, -0.22038656), rtol=1e-6)
    np.testing.assert_allclose(ab(12), (0.64911213, -0.35160898), rtol=1e-6)
    np.testing.assert_allclose(ab(24), (0.999996, -0.002), rtol=1e-5)  # SURVEY quotes 4 significant digits here
    np.testing.assert_allclose(1 / math.sqrt(sig[0] ** 2 + 1), 1.4285699e-3, rtol=1e-6)
    # and the oracle's step() follows that closed form
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, generator=g) * 700
    v = torch.randn(2, 3, generator=g)
    a, b = ab(0)
    got = s.step(v, s.timesteps[0], x)
    np.testing.assert_allclose(got.numpy(), (a * x + b * v).numpy(), rtol=2e-5, atol=1e-3)


# ---------------------------------------------------------------------------------------------------------------
# architecture bookkeeping
# ---------------------------------------------------------------------------------------------------------------
def test_parameter_counts_match_svd():
    from posetraj_b200.config import SVDConfig, controlnet_param_shapes, unet_param_shapes
    cfg = SVDConfig()
    n_unet = sum(math.prod(s) for s in unet_param_shapes(cfg).values())
    n_cnet = sum(math.prod(s) for s in controlnet_param_shapes(cfg).values())
    assert round(n_unet / 1e6, 1) == 1524.6          # SVD's published 1.52 B
    assert round(n_cnet / 1e6, 1) == 682.0
    n_cam = sum(math.prod(s) for s in controlnet_param_shapes(cfg, cam=True).values())
    assert n_cam - n_cnet == 268 * 256 + 256          # cc_projection


def test_key_tree_equals_oracle_modules():
    """The product's expected state-dict (config.py) and the oracle's nn.Module tree name the same tensors."""
    from oracle.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.config import SVDConfig, controlnet_param_shapes, unet_param_shapes
    cfg = SVDConfig()
    with torch.device("meta"):
        unet = UNetSpatioTemporalConditionControlNetModel()
        cnet = ControlNetSDVModel(cam=True, bbox=True)
    for module, shapes in ((unet, unet_param_shapes(cfg)), (cnet, controlnet_param_shapes(cfg, cam=True, bbox=True))):
        sd = {k: tuple(v.shape) for k, v in module.state_dict().items()}
        assert set(sd) == set(shapes), (sorted(set(sd) ^ set(shapes))[:6])
        assert all(tuple(shapes[k]) == sd[k] for k in sd)


def test_residual_multipliers_are_the_loop_accumulation():
    """Simulate the reference's in-loop add (unet...controlnet.py:451-459) symbolically."""
    from posetraj_b200.config import SVDConfig, residual_multipliers
    cfg = SVDConfig()
    per_block = [3, 3, 3, 2]
    skips = [0]            # conv_in output; each entry counts how often its residual was added
    for n_new in per_block:
        skips += [0] * n_new
        skips = [c + 1 for c in skips]      # zip(skips, residuals): every skip collected so far gets +r_i
    assert skips == [4, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1]
    assert residual_multipliers(cfg) == skips


def test_flop_counter_reproduces_survey_totals():
    from posetraj_b200.config import SVDConfig
    from posetraj_b200.roofline import step_flops
    cfg = SVDConfig()
    t, c = step_flops(cfg)
    assert round(t / 1e12, 2) == 33.34 and round(c["_unet"] / 1e12, 2) == 24.28 and round(c["_controlnet"] / 1e12, 2) == 9.06
    assert round(step_flops(cfg, essential=True)[0] / 1e12, 2) == 31.82
    assert round(step_flops(cfg, frames=25, h=72, w=128)[0] / 1e12, 1) == 220.1
    assert round(step_flops(cfg, cam=True)[1]["_controlnet"] / 1e12, 3) == 9.075


def test_sinusoidal_embedding_known_answer():
    from oracle.svd_blocks import sinusoidal_embedding
    e = sinusoidal_embedding(torch.tensor([0.0, 1.0]), 8)
    assert torch.allclose(e[0], torch.tensor([1, 1, 1, 1, 0, 0, 0, 0.0]))      # cos first (flip_sin_to_cos)
    f = torch.exp(-math.log(10000.0) * torch.arange(4) / 4)
    assert torch.allclose(e[1], torch.cat([torch.cos(f), torch.sin(f)]))


# ---------------------------------------------------------------------------------------------------------------
# invariants of SURVEY.md §4 on the small SVD-shaped config
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def small():
    cfg = small_cfg()
    inp = make_small_inputs(cfg, h=8, w=16)
    x = torch.cat([torch.cat([inp["latents"]] * 2) / 700.0, inp["image_latents"]], dim=2)
    t = torch.tensor(1.2)
    return cfg, inp, x, t


def test_faithful_init_gives_zero_residuals_and_plain_unet(small):
    cfg, inp, x, t = small
    unet, cnet = oracle_pair(cfg, seed=3, randomize_zero_convs=False)
    with torch.no_grad():
        down, mid = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"])
        assert len(down) == 12
        assert all(float(d.abs().max()) == 0.0 for d in down) and float(mid.abs().max()) == 0.0
        a = unet(x, t, inp["image_embeddings"], down_block_additional_residuals=down,
                 mid_block_additional_residual=mid, added_time_ids=inp["added_time_ids"])
        b = unet(x, t, inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
    assert torch.equal(a, b)


def test_in_loop_accumulation_equals_multiplied_residuals(small):
    """Feeding m_i * r_i once (what the fused kernels do) == the reference's repeated in-loop add."""
    from posetraj_b200.config import residual_multipliers
    cfg, inp, x, t = small
    unet, cnet = oracle_pair(cfg, seed=4)
    with torch.no_grad():
        down, mid = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"])
        want = unet(x, t, inp["image_embeddings"], down_block_additional_residuals=down,
                    mid_block_additional_residual=mid, added_time_ids=inp["added_time_ids"])
        # same network, residuals pre-multiplied, added exactly once (after the loop): patch the loop away
        mult = residual_multipliers(cfg)
        emb = unet.embed(x, t, inp["added_time_ids"])
        b, f = x.shape[:2]
        h = unet.conv_in(x.flatten(0, 1))
        ehs = inp["image_embeddings"].repeat_interleave(f, dim=0)
        ioi = torch.zeros(b, f)
        skips = (h,)
        for blk in unet.down_blocks:
            h, res = blk(h, emb, ehs, ioi) if blk.has_cross_attention else blk(h, emb, ioi)
            skips += res
        skips = tuple(s + m * r for s, m, r in zip(skips, mult, down))
        h = unet.mid_block(h, emb, ehs, ioi) + mid
        for blk in unet.up_blocks:
            k = len(blk.resnets)
            res, skips = skips[-k:], skips[:-k]
            h = blk(h, res, emb, ehs, ioi) if blk.has_cross_attention else blk(h, res, emb, ioi)
        got = unet.conv_out(unet.conv_act(unet.conv_norm_out(h))).reshape(b, f, *want.shape[2:])
    assert rel_l2(got, want) < 1e-5


def test_batch_rows_depend_on_other_rows_embeddings_not_latents(small):
    """the CFG split is legal iff every rank holds all rows' image embeddings."""
    cfg, inp, x, t = small
    unet, _ = oracle_pair(cfg, seed=5)
    with torch.no_grad():
        base = unet(x, t, inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
        x2 = x.clone()
        x2[1] += 0.5                                      # change row 1's latents only
        a = unet(x2, t, inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
        assert torch.equal(a[0], base[0])                 # row 0 untouched
        e2 = inp["image_embeddings"].clone()
        e2[1] += 0.5                                      # change row 1's embedding only
        b = unet(x, t, e2, added_time_ids=inp["added_time_ids"])
        assert not torch.equal(b[0], base[0])             # row 0 DOES change: the temporal context interleave


def test_one_token_cross_attention_is_a_constant_vector():
    """softmax over one key is 1, so attn2(x, e) = to_out(to_v(e)) for every query."""
    from oracle.svd_blocks import Attention
    torch.manual_seed(0)
    attn = Attention(128, heads=2, dim_head=64, cross_attention_dim=64)
    x = torch.randn(3, 10, 128)
    e = torch.randn(3, 1, 64)
    with torch.no_grad():
        want = attn(x, e)
        vec = attn.to_out[0](attn.to_v(e))
    assert torch.allclose(want, vec.expand_as(want), atol=1e-6)


def test_camera_branch_identity_init_is_noop_iff_camera_is_zero(small):
    cfg, inp, x, t = small
    _, cnet = oracle_pair(cfg, seed=6, cam=True)
    kw = dict(controlnet_cond=inp["controlnet_condition"])
    with torch.no_grad():
        plain = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], **kw)
        zero = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], camera_cond=torch.zeros(2, cfg.num_frames, 12), **kw)
        moved = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], camera_cond=inp["camera_cond"], **kw)
    # oracle_pair rounds weights to bf16 values; the identity block stays exactly the identity
    assert all(rel_l2(a, b) < 1e-6 for a, b in zip(zero[0], plain[0]))
    assert any(rel_l2(a, b) > 1e-4 for a, b in zip(moved[0], plain[0]))


def test_bbox_tower_reuses_conv_out(small):
    """controlnet_sdv_bbox.py:134: tower 2 is projected with the shared conv_out; conv_out_2 is dead."""
    cfg, inp, x, t = small
    _, cnet = oracle_pair(cfg, seed=7, bbox=True)
    with torch.no_grad():
        a = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"],
                 controlnet_bbox=inp["controlnet_condition"].flip(-1))
        torch.nn.init.normal_(cnet.controlnet_cond_embedding.conv_out_2.weight)
        b = cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"],
                 controlnet_bbox=inp["controlnet_condition"].flip(-1))
    assert all(torch.equal(p, q) for p, q in zip(a[0], b[0]))


def test_cpu_sample_plan_is_bounded():
    import bench
    assert bench._cpu_sample_plan(100.0) == (2, 40, 72)
    f, h, w = bench._cpu_sample_plan(0.5)
    assert f == 1 and h % 8 == 0 and w % 8 == 0


def test_product_refuses_cpu():
    """No CPU fallback: constructing a network mirror without CUDA raises."""
    from posetraj_b200.config import unet_param_shapes
    from posetraj_b200.models import UNetSpatioTemporalConditionControlNetModel
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cfg = small_cfg()
    with pytest.raises(RuntimeError):
        UNetSpatioTemporalConditionControlNetModel(cfg, {k: torch.zeros(s) for k, s in unet_param_shapes(cfg).items()})


def test_oracle_reproduces_committed_model_golden():
    """tests/golden/model_golden.safetensors (gen_model_golden.py): one denoise step of the small config, cam model."""
    from safetensors.torch import load_file
    g = load_file(str(GOLDEN / "model_golden.safetensors"))
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=0, cam=True)
    inp = make_small_inputs(cfg)
    with torch.no_grad():
        down, mid = o_cnet(g["sample"], g["timestep"][0], inp["image_embeddings"], inp["added_time_ids"],
                           controlnet_cond=inp["controlnet_condition"], camera_cond=inp["camera_cond"], conditioning_scale=0.8)
        pred = o_unet(g["sample"], g["timestep"][0], inp["image_embeddings"], down_block_additional_residuals=down,
                      mid_block_additional_residual=mid, added_time_ids=inp["added_time_ids"])
    assert rel_l2(pred, g["noise_pred"]) < 1e-5       # same code, same seeds: only thread-count-dependent summation order
    assert rel_l2(mid, g["mid_residual"]) < 1e-5 and rel_l2(down[11], g["down_residual_11"]) < 1e-5
