"""PNDM (PLMS) scheduler for the denoising loop — host-side tables + one fused device step.

Restates diffusers 0.9 `PNDMScheduler(skip_prk_steps=True, steps_offset=1, beta_schedule="scaled_linear",
beta_start=0.00085, beta_end=0.012, set_alpha_to_one=False)` [memory; SURVEY Appendix B], i.e. the scheduler
`RiffusionPipeline.interpolate_img2img` drives at riffusion/riffusion_pipeline.py:314,361-365,379,392-396,403,418.
The scalar recurrences (alphas, timestep table, multistep weights) run on the host in fp32; the tensor update
runs in one kernel fused with the classifier-free-guidance combine (`rf_cfg_pndm_step_f16`).
"""
from __future__ import annotations

import types
import typing as T

import numpy as np
import torch

from riffusion import tc_ops as ops


class PNDMSchedulerB200:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.config = {"steps_offset": steps_offset, "num_train_timesteps": num_train_timesteps}
        self.init_noise_sigma = 1.0
        self.timesteps: T.Optional[torch.Tensor] = None
        self.set_timesteps(50)

    # -- schedule -------------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        base = (np.arange(0, num_inference_steps) * ratio).round() + self.config["steps_offset"]
        plms = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy()      # 961 duplicated
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets: T.List[torch.Tensor] = []
        self.counter = 0
        self.cur_sample: T.Optional[torch.Tensor] = None

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def _alpha(self, t: int) -> float:
        return float(self.alphas_cumprod[t]) if t >= 0 else float(self.final_alpha_cumprod)

    def coefficients(self, timestep: int, prev_timestep: int) -> T.Tuple[float, float]:
        a_t, a_p = self._alpha(timestep), self._alpha(prev_timestep)
        b_t, b_p = 1.0 - a_t, 1.0 - a_p
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return (a_p / a_t) ** 0.5, (a_p - a_t) / denom

    def add_noise(self, original: torch.Tensor, noise: torch.Tensor, timestep, mask=None, blend_with=None) -> torch.Tensor:
        a = float(self.alphas_cumprod[int(timestep)])
        return ops.axpby(original.contiguous(), noise.contiguous(), a ** 0.5, (1.0 - a) ** 0.5, mask, blend_with)

    # -- one multistep update ------------------------------------------------------------------------
    def plan(self, timestep: int):
        """Host bookkeeping of one PLMS step: returns (coef4, history tensors, sample_override, push, ca, cb)."""
        timestep = int(timestep)
        prev = timestep - self.num_train_timesteps // self.num_inference_steps
        push = self.counter != 1
        if not push:                                   # 2nd call: redo the first step from the saved sample
            prev, timestep = timestep, timestep + self.num_train_timesteps // self.num_inference_steps
        n_hist = len(self.ets[-3:]) + 1 if push else len(self.ets)
        override = None
        if n_hist == 1 and self.counter == 0:
            coef, hist = (1.0, 0.0, 0.0, 0.0), []
        elif n_hist == 1 and self.counter == 1:
            coef, hist, override = (0.5, 0.5, 0.0, 0.0), [self.ets[-1]], self.cur_sample
        elif n_hist == 2:
            coef, hist = (1.5, -0.5, 0.0, 0.0), [self.ets[-1]]
        elif n_hist == 3:
            coef, hist = (23 / 12, -16 / 12, 5 / 12, 0.0), [self.ets[-1], self.ets[-2]]
        else:
            coef, hist = (55 / 24, -59 / 24, 37 / 24, -9 / 24), [self.ets[-1], self.ets[-2], self.ets[-3]]
        ca, cb = self.coefficients(timestep, prev)
        return coef, hist, override, push, ca, cb

    def step_cfg(self, eps_pair: torch.Tensor, guidance: float, timestep: int, sample: torch.Tensor) -> torch.Tensor:
        """Guidance combine (riffusion_pipeline.py:411-415) + scheduler.step (:418) in one kernel.
        eps_pair = UNet output for [uncond | text]."""
        coef, hist, override, push, ca, cb = self.plan(timestep)
        if self.counter == 0:
            self.cur_sample = sample
        base = sample if override is None else override
        eps, prev = ops.cfg_pndm_step(eps_pair.contiguous(), guidance, hist, coef, base.contiguous(), ca, cb, want_eps=push)
        if push:
            self.ets = self.ets[-3:] + [eps]
        elif override is not None:
            self.cur_sample = None
        self.counter += 1
        return prev

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, **kwargs):
        """diffusers-compatible signature: the model output is already guided."""
        pair = torch.cat([model_output, model_output]).contiguous()     # eps_u == eps_t  =>  guided eps == eps
        return types.SimpleNamespace(prev_sample=self.step_cfg(pair, 0.0, int(timestep), sample))
