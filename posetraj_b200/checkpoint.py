"""Loading diffusers-format checkpoints into the kernel-backed mirrors.

The reference loads its networks with diffusers' `ModelMixin.from_pretrained(path, subfolder=...)`
(/root/reference/scripts/run_inference_vipseg_json_repro.py:335-337): a directory holding `config.json` and
`diffusion_pytorch_model[.fp16].safetensors` (or `.bin`).  The state-dict key tree of our mirrors IS the diffusers
one (posetraj_b200/config.py lists every key and shape and the constructors verify them), so loading is reading the
file and, if present, the architecture values of config.json.  No network access, no diffusers import.
"""
from __future__ import annotations

import json
import os
from dataclasses import fields
from typing import Dict, Optional, Tuple

import torch

from .config import SVDConfig

_WEIGHT_NAMES = ("diffusion_pytorch_model{v}.safetensors", "diffusion_pytorch_model{v}.bin")


def resolve_dir(path: str, subfolder: Optional[str] = None) -> str:
    d = os.path.join(path, subfolder) if subfolder else path
    if not os.path.isdir(d):
        raise FileNotFoundError(f"checkpoint directory not found: {d}")
    return d


def load_config(directory: str, **overrides) -> SVDConfig:
    """SVDConfig from `config.json` (keys our config does not know are ignored, e.g. `_class_name`, `down_block_types`);
    missing file -> the SVD img2vid defaults."""
    known = {f.name for f in fields(SVDConfig)}
    kw: Dict = {}
    cfg_path = os.path.join(directory, "config.json")
    if os.path.exists(cfg_path):
        with open(cfg_path) as f:
            raw = json.load(f)
        for k, v in raw.items():
            if k in known and v is not None:
                kw[k] = tuple(v) if isinstance(v, list) else v
    kw.update({k: v for k, v in overrides.items() if k in known})
    return SVDConfig(**kw)


def load_state_dict(directory: str, variant: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Reads `diffusion_pytorch_model[.variant].safetensors` (preferred) or `.bin` from `directory` (CPU tensors)."""
    v = f".{variant}" if variant else ""
    for pat in _WEIGHT_NAMES:
        p = os.path.join(directory, pat.format(v=v))
        if os.path.exists(p):
            if p.endswith(".safetensors"):
                from safetensors.torch import load_file
                return load_file(p, device="cpu")
            return torch.load(p, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no diffusion_pytorch_model{v}.safetensors / .bin in {directory}")


def save_pretrained(state_dict: Dict[str, torch.Tensor], cfg: SVDConfig, directory: str, class_name: str,
                    variant: Optional[str] = None) -> str:
    """Writes the diffusers layout (used by the tests to mint checkpoints; mirrors `ModelMixin.save_pretrained`)."""
    from dataclasses import asdict
    from safetensors.torch import save_file
    os.makedirs(directory, exist_ok=True)
    meta = {k: (list(v) if isinstance(v, tuple) else v) for k, v in asdict(cfg).items()}
    meta["_class_name"] = class_name
    with open(os.path.join(directory, "config.json"), "w") as f:
        json.dump(meta, f, indent=1)
    v = f".{variant}" if variant else ""
    path = os.path.join(directory, f"diffusion_pytorch_model{v}.safetensors")
    save_file({k: t.detach().cpu().contiguous() for k, t in state_dict.items()}, path)
    return path


def detect_controlnet_flags(state_dict: Dict[str, torch.Tensor]) -> Tuple[bool, bool]:
    """(cam, bbox): `cc_projection` marks models/controlnet_sdv_cam_infer.py checkpoints, `conv_in_2` the bbox tower."""
    cam = "controlnet_cond_embedding.cc_projection.weight" in state_dict
    bbox = "controlnet_cond_embedding.conv_in_2.weight" in state_dict
    return cam, bbox
