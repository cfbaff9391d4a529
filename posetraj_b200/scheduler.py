"""EulerDiscreteScheduler mirror (Karras sigmas, continuous v-prediction timesteps) for the SVD config.

Reference: /root/reference/utils/scheduling_euler_discrete_karras_fix.py
  set_timesteps :290-350 (+ _convert_to_karras :376-399), init_noise_sigma :249-255,
  scale_model_input :264-288, step :418-528 (v_prediction branch :504-517).
The sigma table is computed on the host in float64 exactly like the reference (numpy) and kept on the device
so the fused CFG+Euler kernel (csrc/sched.cu) indexes it by a device-side step counter: no `.item()` sync and
no per-step host arithmetic.  `step()` keeps the reference signature and runs the same kernel.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import ops


@dataclass
class EulerDiscreteSchedulerOutput:
    prev_sample: torch.FloatTensor
    pred_original_sample: Optional[torch.FloatTensor] = None


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "v_prediction",
                 interpolation_type: str = "linear", use_karras_sigmas: bool = True, sigma_min: float = 0.002,
                 sigma_max: float = 700.0, timestep_spacing: str = "leading", timestep_type: str = "continuous",
                 steps_offset: int = 1):
        if prediction_type != "v_prediction" or not use_karras_sigmas or timestep_type != "continuous":
            raise ValueError("posetraj_b200 implements the SVD scheduler config: v_prediction, Karras sigmas, "
                             "continuous timesteps")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, prediction_type=prediction_type,
                                      interpolation_type=interpolation_type, use_karras_sigmas=use_karras_sigmas,
                                      sigma_min=sigma_min, sigma_max=sigma_max, timestep_spacing=timestep_spacing,
                                      timestep_type=timestep_type, steps_offset=steps_offset)
        self.sigmas = None
        self.timesteps = None
        self.num_inference_steps = None
        self._step_index = None
        self.step_index_dev = None
        self.is_scale_input_called = False

    # ---- schedule ---------------------------------------------------------------------------------
    @staticmethod
    def karras_sigmas(n: int, sigma_min: float, sigma_max: float) -> np.ndarray:
        rho = 7.0
        ramp = np.linspace(0, 1, n)
        lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
        return (hi + ramp * (lo - hi)) ** rho

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device] = None):
        self.num_inference_steps = num_inference_steps
        sig = self.karras_sigmas(num_inference_steps, self.config.sigma_min, self.config.sigma_max)
        sig = torch.from_numpy(sig).to(dtype=torch.float32)
        self.timesteps = torch.Tensor([0.25 * s.log() for s in sig]).to(device=device)
        self.sigmas = torch.cat([sig, torch.zeros(1)]).to(device=device)
        self._step_index = None
        self.is_scale_input_called = False
        if device is not None and torch.device(device).type == "cuda":
            self.step_index_dev = torch.zeros(1, device=device, dtype=torch.int32)

    @property
    def init_noise_sigma(self):
        m = self.sigmas.max()
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return m
        return (m ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        cand = (self.timesteps == timestep).nonzero()
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    # ---- per-step API (reference signatures) -----------------------------------------------------------
    def scale_model_input(self, sample: torch.FloatTensor, timestep) -> torch.FloatTensor:
        """sample / sqrt(sigma^2 + 1).  Kept for API compatibility; the pipeline's fused kernel produces the scaled
        model input itself and does not call this."""
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        self.is_scale_input_called = True
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output: torch.FloatTensor, timestep, sample: torch.FloatTensor, s_churn: float = 0.0,
             s_tmin: float = 0.0, s_tmax: float = float("inf"), s_noise: float = 1.0, generator=None,
             return_dict: bool = True) -> Union[EulerDiscreteSchedulerOutput, Tuple]:
        if isinstance(timestep, int) or isinstance(timestep, (torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                             " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                             " one of the `scheduler.timesteps` as a timestep.")
        if s_churn != 0.0:
            raise ValueError("posetraj_b200: s_churn > 0 (stochastic sampling) is not on the PoseTraj path")
        if model_output.device.type != "cuda":
            raise RuntimeError("posetraj_b200: scheduler.step runs on CUDA only (no CPU fallback)")
        if self._step_index is None:
            self._init_step_index(timestep)
        shape = sample.shape
        Fr, Cc, H, W = shape[-4:]
        lead = int(np.prod(shape[:-4])) if len(shape) > 4 else 1
        lat = sample.detach().to(torch.float32).reshape(lead * Fr, Cc, H, W).contiguous().clone()
        pred = model_output.detach().to(torch.float32).reshape(lead * Fr, Cc, H, W).contiguous()
        idx = torch.full((1,), self._step_index, device=lat.device, dtype=torch.int32)
        ones = torch.ones(lead * Fr, device=lat.device, dtype=torch.float32)
        ops.CfgEuler(noise_pred=pred, latents=lat, guidance=ones, sigmas=self.sigmas.to(lat.device), step_index=idx,
                     pred_nchw_f32=True, single_pred=True).launch(torch.cuda.current_stream().cuda_stream)
        self._step_index += 1
        prev = lat.reshape(shape).to(model_output.dtype)
        if not return_dict:
            return (prev,)
        return EulerDiscreteSchedulerOutput(prev_sample=prev)

    def __len__(self):
        return self.config.num_train_timesteps
