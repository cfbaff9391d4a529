"""Checkpoint loading: diffusers directory layout -> the kernel-backed mirrors."""
import json

import pytest
import torch

from parity_util import make_small_inputs, oracle_pair, rel_l2, small_cfg


def test_roundtrip_layout_and_config(tmp_path):
    from posetraj_b200 import checkpoint
    cfg = small_cfg()
    _, o_cnet = oracle_pair(cfg, seed=2, cam=True)
    d = tmp_path / "ckpt" / "controlnet"
    path = checkpoint.save_pretrained(o_cnet.state_dict(), cfg, str(d), "ControlNetSDVModel")
    assert path.endswith("diffusion_pytorch_model.safetensors")
    raw = json.loads((d / "config.json").read_text())
    raw["down_block_types"] = ["CrossAttnDownBlockSpatioTemporal"] * 3 + ["DownBlockSpatioTemporal"]   # unknown keys are ignored
    (d / "config.json").write_text(json.dumps(raw))
    got = checkpoint.load_config(checkpoint.resolve_dir(str(tmp_path / "ckpt"), "controlnet"))
    assert got == cfg
    sd = checkpoint.load_state_dict(str(d))
    assert set(sd) == set(o_cnet.state_dict()) and all(torch.equal(sd[k], v) for k, v in o_cnet.state_dict().items())
    assert checkpoint.detect_controlnet_flags(sd) == (True, False)
    with pytest.raises(FileNotFoundError):
        checkpoint.load_state_dict(str(d), variant="fp16")
    with pytest.raises(FileNotFoundError):
        checkpoint.resolve_dir(str(tmp_path / "ckpt"), "unet")


def test_missing_config_gives_svd_defaults(tmp_path):
    from posetraj_b200 import checkpoint
    from posetraj_b200.config import SVDConfig
    assert checkpoint.load_config(str(tmp_path)) == SVDConfig()
    assert checkpoint.load_config(str(tmp_path), num_frames=25).num_frames == 25


@pytest.mark.gpu
def test_from_pretrained_runs_like_the_oracle(tmp_path, cuda_dev):
    from posetraj_b200 import checkpoint
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=3)
    checkpoint.save_pretrained(o_unet.state_dict(), cfg, str(tmp_path / "svd" / "unet"), "UNetSpatioTemporalConditionControlNetModel", variant="fp16")
    checkpoint.save_pretrained(o_cnet.state_dict(), cfg, str(tmp_path / "pt" / "controlnet"), "ControlNetSDVModel")
    unet = UNetSpatioTemporalConditionControlNetModel.from_pretrained(str(tmp_path / "svd"), subfolder="unet", variant="fp16", device=cuda_dev)
    cnet = ControlNetSDVModel.from_pretrained(str(tmp_path / "pt"), subfolder="controlnet", device=cuda_dev)
    inp = make_small_inputs(cfg)
    x = torch.cat([torch.cat([inp["latents"]] * 2) / 10.05, inp["image_latents"]], dim=2)
    t = torch.tensor(0.577)
    with torch.no_grad():
        o_down, o_mid = o_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"], controlnet_cond=inp["controlnet_condition"])
        want = o_unet(x, t, inp["image_embeddings"], down_block_additional_residuals=o_down,
                      mid_block_additional_residual=o_mid, added_time_ids=inp["added_time_ids"])
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    down, mid = cnet(x.to(cuda_dev), t.to(cuda_dev), d["image_embeddings"], d["added_time_ids"],
                     controlnet_cond=d["controlnet_condition"], return_dict=False)
    got = unet(x.to(cuda_dev), t.to(cuda_dev), d["image_embeddings"], down_block_additional_residuals=down,
               mid_block_additional_residual=mid, added_time_ids=d["added_time_ids"]).sample
    assert rel_l2(got, want) < 1e-2
    # a checkpoint with a missing tensor is rejected, not silently zero-filled
    sd = checkpoint.load_state_dict(str(tmp_path / "pt" / "controlnet"))
    sd.pop("conv_in.weight")
    with pytest.raises(KeyError):
        ControlNetSDVModel(cfg, sd, cuda_dev)


@pytest.mark.gpu
def test_pipeline_from_pretrained_like_the_reference_script(tmp_path, cuda_dev):
    """scripts/run_inference_vipseg_json_repro.py:335-339, line for line: ControlNet / UNet / pipeline `from_pretrained`,
    then `enable_model_cpu_offload()` (a no-op here), then one call."""
    import json
    from posetraj_b200 import checkpoint
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    cfg = small_cfg()
    o_unet, o_cnet = oracle_pair(cfg, seed=3)
    root = tmp_path / "stable-video-diffusion-img2vid"
    checkpoint.save_pretrained(o_unet.state_dict(), cfg, str(root / "unet"), "UNetSpatioTemporalConditionControlNetModel")
    checkpoint.save_pretrained(o_cnet.state_dict(), cfg, str(tmp_path / "ckpt" / "controlnet"), "ControlNetSDVModel")
    (root / "scheduler").mkdir(parents=True)
    (root / "scheduler" / "scheduler_config.json").write_text(json.dumps(
        {"_class_name": "EulerDiscreteScheduler", "sigma_min": 0.002, "sigma_max": 700.0, "use_karras_sigmas": True,
         "prediction_type": "v_prediction", "timestep_type": "continuous", "beta_schedule": "scaled_linear"}))
    controlnet = ControlNetSDVModel.from_pretrained(str(tmp_path / "ckpt"), subfolder="controlnet", device=cuda_dev)
    unet = UNetSpatioTemporalConditionControlNetModel.from_pretrained(str(root), subfolder="unet", device=cuda_dev)
    pipeline = StableVideoDiffusionPipelineControlNet.from_pretrained(str(root), controlnet=controlnet, unet=unet)
    pipeline.enable_model_cpu_offload()
    assert pipeline.vae is None and pipeline.image_encoder is None      # not in this directory: passed per call instead
    inp = make_small_inputs(cfg)
    h, w = inp["latents"].shape[-2:]
    out = pipeline(None, inp["controlnet_condition"][0].to(cuda_dev), height=h * 8, width=w * 8, num_frames=cfg.num_frames,
                   num_inference_steps=2, latents=(inp["latents"] / 700.0).to(cuda_dev), output_type="latent",
                   image_embeddings=inp["image_embeddings"].to(cuda_dev), image_latents=inp["image_latents"].to(cuda_dev)).frames
    assert out.shape == (1, cfg.num_frames, 4, h, w) and torch.isfinite(out).all()
    with pytest.raises(FileNotFoundError):
        StableVideoDiffusionPipelineControlNet.from_pretrained(str(tmp_path / "nope"), controlnet=controlnet, unet=unet)
