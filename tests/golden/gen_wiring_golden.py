"""Pins the WIRING half of the oracle against the reference's own code ("next round" item 1b).

Executes, unmodified, in this container:
  /root/reference/models/controlnet_sdv.py            ControlNetSDVModel.__init__ / forward          (plain)
  /root/reference/models/controlnet_sdv_cam_infer.py  ... with ControlNetConditioningEmbeddingSVD_CAM (camera branch)
  /root/reference/models/controlnet_sdv_bbox.py       ... with the second (bbox) tower
  /root/reference/models/unet_spatio_temporal_condition_controlnet.py   UNet __init__ / forward (residual injection)
  /root/reference/models/modified_svd.py              forward_TemporalBasicTransformerBlock (:50-114),
                                                      forward_TransformerSpatioTemporalModel (:118-223),
                                                      forward_CrossAttn{Up,Down}BlockSpatioTemporal (:225-348)
`diffusers` (absent from this image) is replaced by stand-ins in sys.modules: plumbing classes (ModelMixin, ConfigMixin,
register_to_config, BaseOutput, ...) and, for `diffusers.models.unet_3d_blocks` / `.embeddings`, the LEAF blocks of
oracle/svd_blocks.py (ResnetBlock2D, TemporalResnetBlock, Attention, FeedForward, BasicTransformerBlock, AlphaBlender,
samplers).  Every forward the reference repository itself carries is run from the reference's file: the four functions
of modified_svd.py are bound as the `forward` of the corresponding block classes.  What is pinned: constructor wiring
(channel plans, eps, zero-conv order), the in-loop residual accumulation (multipliers [4,4,4,4,3,3,3,2,2,2,1,1]), the
temporal-context interleave, the camera / bbox branches, conditioning_scale.  What stays unpinned: the
arithmetic of the leaf blocks (diffusers 0.24.0 restated from SURVEY Appendix A).

Outputs (small SVD-shaped config, parity_util.small_cfg, 3 frames, 16x24 latent) go to tests/golden/wiring_golden.safetensors:
the UNet noise prediction in full, and for each of the 13 ControlNet residuals its fp64 sum, L2 norm and 512 strided
samples, for the plain / camera / bbox variants.  tests/test_wiring_golden_cpu.py asserts that oracle/models.py
reproduces them.  Run here only (needs /root/reference):  python tests/golden/gen_wiring_golden.py
"""
import functools
import importlib.util
import inspect
import math
import os
import sys
import types
from types import SimpleNamespace

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wiring_golden.safetensors")
N_SAMPLES = 512


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def install_shims():
    """diffusers stand-ins: plumbing + oracle leaf blocks; returns the module holding the reference's block forwards."""
    from oracle import svd_blocks as ob

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *a, **kw):
            bound = inspect.signature(init).bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            init(self, *a, **kw)
            object.__setattr__(self, "config", SimpleNamespace(**cfg))
        return wrapper

    class ConfigMixin:
        pass

    class ModelMixin(nn.Module):
        pass

    class BaseOutput:
        pass

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    logging = SimpleNamespace(get_logger=lambda name: SimpleNamespace(warning=lambda *a, **k: None,
                                                                      info=lambda *a, **k: None))
    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    mod("diffusers.loaders", FromOriginalControlnetMixin=type("FromOriginalControlnetMixin", (), {}),
        UNet2DConditionLoadersMixin=type("UNet2DConditionLoadersMixin", (), {}))
    mod("diffusers.utils", BaseOutput=BaseOutput, logging=logging, is_torch_version=lambda op, v: True)
    mod("diffusers.models", UNetSpatioTemporalConditionModel=_Dummy)
    mod("diffusers.models.attention_processor", ADDED_KV_ATTENTION_PROCESSORS=(), CROSS_ATTENTION_PROCESSORS=(),
        AttentionProcessor=_Dummy, AttnAddedKVProcessor=_Dummy, AttnProcessor=_Dummy)
    mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)

    class Timesteps(ob.Timesteps):   # diffusers signature: Timesteps(num_channels, flip_sin_to_cos, downscale_freq_shift)
        def __init__(self, num_channels, flip_sin_to_cos=True, downscale_freq_shift=0):
            assert flip_sin_to_cos is True and downscale_freq_shift == 0
            super().__init__(num_channels)

    mod("diffusers.models.embeddings", TextImageProjection=_Dummy, TextImageTimeEmbedding=_Dummy, TextTimeEmbedding=_Dummy,
        TimestepEmbedding=ob.TimestepEmbedding, Timesteps=Timesteps)

    # ---- the reference's own block forwards (modified_svd.py) bound onto the oracle's block containers --------------
    ref_fw = _load("ref_modified_svd", os.path.join(REF, "modified_svd.py"))

    class RefAlphaBlender(ob.AlphaBlender):
        def forward(self, x_spatial, x_temporal, image_only_indicator, camera_para=None):
            assert camera_para is None
            return super().forward(x_spatial, x_temporal, image_only_indicator)

    class RefTemporalBlock(ob.TemporalBasicTransformerBlock):
        _chunk_size, _chunk_dim, is_res = None, 0, True
        forward = ref_fw.forward_TemporalBasicTransformerBlock            # modified_svd.py:50-114

    class RefSTModel(ob.TransformerSpatioTemporalModel):
        gradient_checkpointing = False
        forward = ref_fw.forward_TransformerSpatioTemporalModel           # modified_svd.py:118-223

    class RefCrossDown(ob.CrossAttnDownBlockSpatioTemporal):
        gradient_checkpointing = False
        forward = ref_fw.forward_CrossAttnDownBlockSpatioTemporal         # modified_svd.py:287-348

    class RefCrossUp(ob.CrossAttnUpBlockSpatioTemporal):
        gradient_checkpointing = False
        forward = ref_fw.forward_CrossAttnUpBlockSpatioTemporal           # modified_svd.py:225-285

    # blocks whose forward only exists in diffusers: the oracle's, behind the keyword names the reference calls them with
    class Down(ob.DownBlockSpatioTemporal):
        def forward(self, hidden_states, temb=None, image_only_indicator=None):
            return super().forward(hidden_states, temb, image_only_indicator)

    class Up(ob.UpBlockSpatioTemporal):
        def forward(self, hidden_states, res_hidden_states_tuple, temb=None, image_only_indicator=None):
            return super().forward(hidden_states, res_hidden_states_tuple, temb, image_only_indicator)

    class Mid(ob.UNetMidBlockSpatioTemporal):
        def __init__(self, in_channels, temb_channels, transformer_layers_per_block=1, cross_attention_dim=1280,
                     num_attention_heads=1):
            assert transformer_layers_per_block == 1
            super().__init__(in_channels, temb_channels, num_attention_heads, cross_attention_dim)
            rebind(self)

        def forward(self, hidden_states, temb=None, encoder_hidden_states=None, image_only_indicator=None):
            # diffusers UNetMidBlockSpatioTemporal.forward (no copy in the reference repo): oracle restatement; the
            # attention inside runs the reference's forward_TransformerSpatioTemporalModel and returns a tuple
            h = self.resnets[0](hidden_states, temb, image_only_indicator=image_only_indicator)
            for attn, resnet in zip(self.attentions, self.resnets[1:]):
                h = attn(h, encoder_hidden_states=encoder_hidden_states, image_only_indicator=image_only_indicator,
                         return_dict=False)[0]
                h = resnet(h, temb, image_only_indicator=image_only_indicator)
            return h

    def rebind(block):
        for m in block.modules():
            if type(m) is ob.TransformerSpatioTemporalModel:
                m.__class__ = RefSTModel
                m.time_mixer.__class__ = RefAlphaBlender
            elif type(m) is ob.TemporalBasicTransformerBlock:
                m.__class__ = RefTemporalBlock
        return block

    def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                       num_attention_heads, resnet_eps=None, resnet_act_fn=None, cross_attention_dim=None,
                       transformer_layers_per_block=1, **kw):
        assert transformer_layers_per_block == 1 and resnet_act_fn == "silu" and not kw, kw
        if down_block_type == "CrossAttnDownBlockSpatioTemporal":
            return rebind(RefCrossDown(in_channels, out_channels, temb_channels, num_attention_heads, cross_attention_dim,
                                       add_downsample, num_layers=num_layers))
        assert down_block_type == "DownBlockSpatioTemporal"
        return Down(in_channels, out_channels, temb_channels, add_downsample, num_layers=num_layers)

    def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels, add_upsample,
                     num_attention_heads, resolution_idx=None, resnet_eps=None, resnet_act_fn=None,
                     cross_attention_dim=None, transformer_layers_per_block=1, **kw):
        assert transformer_layers_per_block == 1 and resnet_act_fn == "silu" and not kw, kw
        if up_block_type == "CrossAttnUpBlockSpatioTemporal":
            return rebind(RefCrossUp(in_channels, prev_output_channel, out_channels, temb_channels, num_attention_heads,
                                     cross_attention_dim, add_upsample, num_layers=num_layers))
        assert up_block_type == "UpBlockSpatioTemporal"
        return Up(in_channels, prev_output_channel, out_channels, temb_channels, add_upsample, num_layers=num_layers)

    mod("diffusers.models.unet_3d_blocks", get_down_block=get_down_block, get_up_block=get_up_block,
        UNetMidBlockSpatioTemporal=Mid)
    return ref_fw


def summarize(t: torch.Tensor):
    flat = t.detach().double().reshape(-1)
    stride = max(1, flat.numel() // N_SAMPLES)
    return torch.tensor([float(flat.sum()), float(flat.norm())], dtype=torch.float64), \
        flat[::stride][:N_SAMPLES].float().contiguous()


def run():
    import contextlib
    import io
    from parity_util import make_small_bbox_maps, make_small_inputs, oracle_pair, small_cfg
    install_shims()
    cfg = small_cfg()
    ckw = dict(in_channels=cfg.in_channels, block_out_channels=cfg.block_out_channels,
               addition_time_embed_dim=cfg.addition_time_embed_dim,
               projection_class_embeddings_input_dim=cfg.projection_class_embeddings_input_dim,
               layers_per_block=cfg.layers_per_block, cross_attention_dim=cfg.cross_attention_dim,
               num_attention_heads=cfg.num_attention_heads, num_frames=cfg.num_frames)
    ref_unet_mod = _load("ref_unet", os.path.join(REF, "unet_spatio_temporal_condition_controlnet.py"))
    variants = {"plain": ("controlnet_sdv.py", dict(cam=False, bbox=False)),
                "cam": ("controlnet_sdv_cam_infer.py", dict(cam=True, bbox=False)),
                "bbox": ("controlnet_sdv_bbox.py", dict(cam=False, bbox=True))}
    inp = make_small_inputs(cfg)
    sigma = 10.0
    x = torch.cat([torch.cat([inp["latents"]] * 2) / (sigma ** 2 + 1) ** 0.5, inp["image_latents"]], dim=2)
    t = torch.tensor(0.25 * math.log(sigma))
    bbox_maps = make_small_bbox_maps(cfg, inp)
    out = {"sample": x, "timestep": t.reshape(1)}
    torch.set_num_threads(4)
    for name, (fname, flags) in variants.items():
        o_unet, o_cnet = oracle_pair(cfg, seed=0, **flags)
        with contextlib.redirect_stdout(io.StringIO()):     # the reference constructors print
            ref_cnet_mod = _load("ref_cnet_" + name, os.path.join(REF, fname))
            r_cnet = ref_cnet_mod.ControlNetSDVModel(**ckw).eval()
            r_unet = ref_unet_mod.UNetSpatioTemporalConditionControlNetModel(out_channels=cfg.out_channels, **ckw).eval()
        r_cnet.load_state_dict(o_cnet.state_dict(), strict=True)     # same key tree as the reference's modules
        r_unet.load_state_dict(o_unet.state_dict(), strict=True)
        kw = {}
        if flags["cam"]:
            kw["camera_cond"] = inp["camera_cond"]
        if flags["bbox"]:
            kw["controlnet_bbox"] = bbox_maps
        with torch.no_grad():
            r_down, r_mid = r_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"],
                                   controlnet_cond=inp["controlnet_condition"], conditioning_scale=0.8, return_dict=False, **kw)
            r_pred = r_unet(x, t, inp["image_embeddings"], down_block_additional_residuals=r_down,
                            mid_block_additional_residual=r_mid, added_time_ids=inp["added_time_ids"], return_dict=False)[0]
            o_down, o_mid = o_cnet(x, t, inp["image_embeddings"], inp["added_time_ids"],
                                   controlnet_cond=inp["controlnet_condition"], conditioning_scale=0.8, **kw)
            o_pred = o_unet(x, t, inp["image_embeddings"], down_block_additional_residuals=o_down,
                            mid_block_additional_residual=o_mid, added_time_ids=inp["added_time_ids"])
        out[f"{name}.noise_pred"] = r_pred.contiguous()
        worst = 0.0
        for i, (r, o) in enumerate(zip(list(r_down) + [r_mid], list(o_down) + [o_mid])):
            st, smp = summarize(r)
            out[f"{name}.res{i}.stats"], out[f"{name}.res{i}.samples"] = st, smp
            worst = max(worst, float((r - o).abs().max()))
        print(f"{name}: oracle vs reference wiring: max |diff| residuals {worst:.3e}, noise_pred "
              f"{float((r_pred - o_pred).abs().max()):.3e} (|pred| max {float(r_pred.abs().max()):.3f})")
    return out


if __name__ == "__main__":
    from safetensors.torch import save_file
    res = run()
    save_file({k: v.contiguous() for k, v in res.items()}, OUT,
              metadata={"config": "parity_util.small_cfg()", "weights": "oracle_pair(seed=0, cam/bbox per variant), bf16-valued",
                        "inputs": "parity_util.make_small_inputs(cfg), sigma 10, conditioning_scale 0.8; bbox maps seed 4321",
                        "produced_by": "reference model files executed with diffusers shimmed (see this script's docstring)"})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
