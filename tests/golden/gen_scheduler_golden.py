"""Generates tests/golden/scheduler_golden.json by EXECUTING the reference's own scheduler file
(/root/reference/utils/scheduling_euler_discrete_karras_fix.py) in this container.

The file imports four things from diffusers (not installed here): ConfigMixin, register_to_config, SchedulerMixin /
KarrasDiffusionSchedulers, randn_tensor, BaseOutput, logging.  They are plumbing, not arithmetic, so this script
injects minimal stand-ins into sys.modules and then imports the reference file unmodified.  The numeric body
(set_timesteps :290-350, _convert_to_karras :376-399, scale_model_input :264-288, step :418-528) runs as shipped.

Run here only (needs /root/reference):  python tests/golden/gen_scheduler_golden.py
"""
import functools
import importlib.util
import inspect
import json
import os
import sys
import types
from types import SimpleNamespace

import torch

REF = "/root/reference/utils/scheduling_euler_discrete_karras_fix.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scheduler_golden.json")


def _install_shims():
    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *a, **kw):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            self.config = SimpleNamespace(**cfg)
            init(self, *a, **kw)
        return wrapper

    class ConfigMixin:
        # diffusers' ConfigMixin lets `self.<config key>` fall through to the registered config (deprecated but
        # live in 0.24.0); the reference's __init__ relies on it (`self.use_karras_sigmas` at :225 precedes :244)
        def __getattr__(self, name):
            cfg = self.__dict__.get("config")
            if cfg is not None and hasattr(cfg, name):
                return getattr(cfg, name)
            raise AttributeError(name)

    class SchedulerMixin:
        pass

    class BaseOutput:
        pass

    class _Enum:
        pass

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, device=device, dtype=dtype)

    mods = {
        "diffusers": types.ModuleType("diffusers"),
        "diffusers.configuration_utils": types.ModuleType("diffusers.configuration_utils"),
        "diffusers.utils": types.ModuleType("diffusers.utils"),
        "diffusers.utils.torch_utils": types.ModuleType("diffusers.utils.torch_utils"),
        "diffusers.schedulers": types.ModuleType("diffusers.schedulers"),
        "diffusers.schedulers.scheduling_utils": types.ModuleType("diffusers.schedulers.scheduling_utils"),
    }
    mods["diffusers.configuration_utils"].ConfigMixin = ConfigMixin
    mods["diffusers.configuration_utils"].register_to_config = register_to_config
    mods["diffusers.utils"].BaseOutput = BaseOutput
    mods["diffusers.utils"].logging = SimpleNamespace(get_logger=lambda name: SimpleNamespace(
        warning=lambda *a, **k: None, info=lambda *a, **k: None))
    mods["diffusers.utils.torch_utils"].randn_tensor = randn_tensor
    mods["diffusers.schedulers.scheduling_utils"].KarrasDiffusionSchedulers = []
    mods["diffusers.schedulers.scheduling_utils"].SchedulerMixin = SchedulerMixin
    sys.modules.update(mods)


def main():
    _install_shims()
    spec = importlib.util.spec_from_file_location("ref_sched", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    # SVD scheduler_config.json
    kw = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
              prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002,
              sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1)
    out = {"generator": "tests/golden/gen_scheduler_golden.py", "reference": REF, "config": kw, "cases": []}
    for n in (25, 4, 50):
        s = ref.EulerDiscreteScheduler(**kw)
        s.set_timesteps(n)
        case = {"num_inference_steps": n, "sigmas": [float(v) for v in s.sigmas],
                "timesteps": [float(v) for v in s.timesteps], "init_noise_sigma": float(s.init_noise_sigma)}
        # a full loop on a seeded sample with a seeded "model output": scale_model_input + step, fp32
        g = torch.Generator().manual_seed(1000 + n)
        x = torch.randn(1, 2, 4, 3, 5, generator=g) * s.init_noise_sigma
        scaled, xs = [], []
        for i, t in enumerate(s.timesteps):
            xin = s.scale_model_input(x, t)
            v = torch.randn(x.shape, generator=g)
            x = s.step(v, t, x).prev_sample
            if i in (0, 1, n // 2, n - 1):
                scaled.append({"i": i, "scale_model_input": xin.flatten().tolist()})
                xs.append({"i": i, "prev_sample": x.flatten().tolist()})
        case["seed"] = 1000 + n
        case["shape"] = [1, 2, 4, 3, 5]
        case["scaled"] = scaled
        case["steps"] = xs
        case["final"] = x.flatten().tolist()
        out["cases"].append(case)
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
