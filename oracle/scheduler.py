"""ORACLE (test infrastructure) — Euler-Karras scheduler restated from
/root/reference/utils/scheduling_euler_discrete_karras_fix.py with the SVD scheduler config:
  set_timesteps  :290-350 (+ _convert_to_karras :376-399)   init_noise_sigma :249-255
  scale_model_input :264-288                                 step (v_prediction) :418-528
Pinned against the reference file itself via tests/golden/scheduler_golden.json.
"""
from __future__ import annotations

import numpy as np
import torch


class EulerKarrasOracle:
    order = 1

    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 700.0, timestep_spacing: str = "leading"):
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.timestep_spacing = timestep_spacing
        self.sigmas = None
        self.timesteps = None
        self._step_index = None

    def set_timesteps(self, n: int, device=None):
        rho = 7.0
        ramp = np.linspace(0, 1, n)
        lo, hi = self.sigma_min ** (1 / rho), self.sigma_max ** (1 / rho)
        sig = (hi + ramp * (lo - hi)) ** rho                       # :391-398 (float64 numpy)
        sig = torch.from_numpy(sig).to(dtype=torch.float32)        # :341
        self.timesteps = torch.Tensor([0.25 * s.log() for s in sig])  # :345 continuous v-prediction timesteps
        self.sigmas = torch.cat([sig, torch.zeros(1)])             # :349
        self._step_index = None
        if device is not None:
            self.sigmas = self.sigmas.to(device)
            self.timesteps = self.timesteps.to(device)

    @property
    def init_noise_sigma(self):
        m = self.sigmas.max()
        if self.timestep_spacing in ("linspace", "trailing"):
            return m
        return (m ** 2 + 1) ** 0.5

    def _init_step_index(self, t):
        cand = (self.timesteps == t).nonzero()
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def scale_model_input(self, sample, t):
        if self._step_index is None:
            self._init_step_index(t)
        sigma = self.sigmas[self._step_index]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, t, sample):
        if self._step_index is None:
            self._init_step_index(t)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        x0 = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        derivative = (sample - x0) / sigma
        dt = self.sigmas[self._step_index + 1] - sigma
        prev = (sample + derivative * dt).to(model_output.dtype)
        self._step_index += 1
        return prev
