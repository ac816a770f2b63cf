"""1-D Gaussian diffusion decoder (reference ``modules/diffusion/ddpm.py:140-374``).

Training loss (:345-374): x_t = q_sample(x0, t, noise) and the masked-L1 against the noise run in the
fused kernels ``ops.qsample`` / ``ops.masked_l1``; the UNet in between stays torch/cuDNN this round.
``t`` and ``noise`` may be supplied by the caller (RNG draws #4 and #5 of the reference forward).
Sampling (:232-326) is the plain torch loop (next row).
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...training_lib.losses import masked_l1_loss, masked_l2_loss
from ...utils.tensormask import TensorMask

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start"])


def _beta_schedule(hp, timesteps: int) -> torch.Tensor:
    kind = hp.beta_schedule.identifier
    if kind == "linear":
        scale = 1000 / timesteps
        return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
    if kind == "scaled_linear":
        b0, b1 = hp.beta_schedule.get("beta_start", 0.0015), hp.beta_schedule.get("beta_end", 0.0195)
        return torch.linspace(b0 ** 0.5, b1 ** 0.5, timesteps, dtype=torch.float64) ** 2
    if kind == "cosine":
        s = hp.beta_schedule.get("s", 0.008)
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
        ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    raise ValueError(f"unknown beta schedule {hp.beta_schedule}")


def _gather(table: torch.Tensor, t: torch.Tensor, ndim: int) -> torch.Tensor:
    return table.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


class GaussianDiffusion1D(nn.Module):
    def __init__(self, model: nn.Module, hp):
        super().__init__()
        self.hp = hp
        self.model = model
        self.objective = hp.get("objective", "pred_noise")
        self.loss_type = hp.get("loss_type", "l1")
        self.clamp_range = hp.get("clamp_range", [-1, 1])
        self.ddim_sampling_eta = hp.get("ddim_sampling_eta", 1.0)
        self.sigma = 1.0
        betas = _beta_schedule(hp, hp.timesteps)
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = hp.get("sampling_timesteps", None) or self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        # same 13 float32 buffers, same names, same order as the reference (checkpoint contract)
        for name, val in (
            ("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", ac_prev),
            ("sqrt_alphas_cumprod", torch.sqrt(ac)), ("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - ac)),
            ("log_one_minus_alphas_cumprod", torch.log(1.0 - ac)),
            ("sqrt_recip_alphas_cumprod", torch.sqrt(1.0 / ac)),
            ("sqrt_recipm1_alphas_cumprod", torch.sqrt(1.0 / ac - 1)),
            ("posterior_variance", post_var),
            ("posterior_log_variance_clipped", torch.log(post_var.clamp(min=1e-20))),
            ("posterior_mean_coef1", betas * torch.sqrt(ac_prev) / (1.0 - ac)),
            ("posterior_mean_coef2", (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac)),
            ("p2_loss_weight", (1 + ac / (1 - ac)) ** -0.0),
        ):
            self.register_buffer(name, val.to(torch.float32))

    # ------------------------------------------------------------------ training
    @property
    def loss_fn(self):
        if self.loss_type == "l1":
            return masked_l1_loss
        if self.loss_type == "l2":
            return masked_l2_loss
        raise ValueError(f"invalid loss type {self.loss_type}")

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (_gather(self.sqrt_alphas_cumprod, t, x_start.dim()) * x_start
                + _gather(self.sqrt_one_minus_alphas_cumprod, t, x_start.dim()) * noise)

    def p_losses(self, x_start: TensorMask, t: torch.Tensor, cond: TensorMask,
                 noise: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        batch_weight = kwargs.pop("loss_batch_weight", None)
        if noise is None:
            noise = torch.randn_like(x_start.value)
        x_t, target = ops.qsample(x_start.value, noise, t, self.sqrt_alphas_cumprod,
                                  self.sqrt_one_minus_alphas_cumprod, x_start.mask)
        model_out = self.model(TensorMask(x_t, x_start.mask), t, cond, **kwargs)
        if self.objective == "pred_noise":
            tgt = TensorMask(target, x_start.mask)
        elif self.objective == "pred_x0":
            tgt = x_start
        else:
            raise ValueError(self.objective)
        return self.loss_fn(TensorMask(model_out.value.float(), model_out.mask), tgt, batch_weight=batch_weight)

    def forward(self, img: TensorMask, cond: TensorMask, t: Optional[torch.Tensor] = None,
                noise: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        if t is None:
            t = torch.randint(0, self.num_timesteps, (img.value.size(0),), device=img.device).long()
        return self.p_losses(img, t, cond, noise=noise, **kwargs)

    # ------------------------------------------------------------------ sampling (torch loop)
    @property
    def is_ddim_sampling(self):
        return self.sampling_timesteps < self.num_timesteps

    def predict_start_from_noise(self, x_t, t, noise):
        return (_gather(self.sqrt_recip_alphas_cumprod, t, x_t.dim()) * x_t
                - _gather(self.sqrt_recipm1_alphas_cumprod, t, x_t.dim()) * noise)

    def predict_noise_from_start(self, x_t, t, x0):
        return ((_gather(self.sqrt_recip_alphas_cumprod, t, x_t.dim()) * x_t - x0)
                / _gather(self.sqrt_recipm1_alphas_cumprod, t, x_t.dim()))

    def q_posterior(self, x_start, x_t, t):
        mean = (_gather(self.posterior_mean_coef1, t, x_t.dim()) * x_start
                + _gather(self.posterior_mean_coef2, t, x_t.dim()) * x_t)
        return (mean, _gather(self.posterior_variance, t, x_t.dim()),
                _gather(self.posterior_log_variance_clipped, t, x_t.dim()))

    def model_predictions(self, x: TensorMask, t: torch.Tensor, cond: TensorMask, **kwargs) -> ModelPrediction:
        out = self.model(x, t, cond, **kwargs)
        out = TensorMask(out.value.float(), out.mask)
        if self.objective == "pred_noise":
            x0 = TensorMask(self.predict_start_from_noise(x.value, t, out.value), out.mask).apply_mask()
            return ModelPrediction(out, x0)
        noise = TensorMask(self.predict_noise_from_start(x.value, t, out.value), out.mask).apply_mask()
        return ModelPrediction(noise, out)

    @torch.no_grad()
    def p_sample(self, x: TensorMask, t: int, cond: TensorMask, **kwargs):
        bt = torch.full((x.value.shape[0],), t, device=x.value.device, dtype=torch.long)
        preds = self.model_predictions(x, bt, cond, **kwargs)
        x0 = preds.pred_x_start.apply_mask().value.clamp_(self.clamp_range[0], self.clamp_range[1])
        mean, _, log_var = self.q_posterior(x_start=x0, x_t=x.value, t=bt)
        noise = torch.randn_like(x.value) * self.sigma if t > 0 else 0.0
        img = mean + (0.5 * log_var).exp() * noise
        return TensorMask(img, preds.pred_x_start.mask).apply_mask(), x0

    @torch.no_grad()
    def p_sample_loop(self, start: TensorMask, cond: TensorMask, **kwargs) -> TensorMask:
        img = start
        stride = self.num_timesteps // self.sampling_timesteps
        for t in reversed(range(0, self.num_timesteps, stride)):
            img, _ = self.p_sample(img, t, cond, **kwargs)
        return img

    @torch.no_grad()
    def ddim_sample(self, start: TensorMask, cond: TensorMask, step_noise=None, **kwargs) -> TensorMask:
        """``step_noise``: optional list of the per-step torch.randn_like draws (one per step but the last) — additive
        API for parity tests; the reference draws them inside the loop (ddpm.py:312)."""
        batch, device = start.value.shape[0], self.betas.device
        step_noise = list(step_noise) if step_noise is not None else None
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        img = start
        for time, time_next in zip(times[:-1], times[1:]):
            tc = torch.full((batch,), time, device=device, dtype=torch.long)
            pred_noise, x0 = self.model_predictions(img, tc, cond, **kwargs)
            x0.value.clamp_(self.clamp_range[0], self.clamp_range[1])
            x0 = x0.apply_mask()
            if time_next < 0:
                img = x0
                continue
            a, a_next = self.alphas_cumprod[time], self.alphas_cumprod[time_next]
            sigma = self.ddim_sampling_eta * ((1 - a / a_next) * (1 - a_next) / (1 - a)).sqrt()
            c = (1 - a_next - sigma ** 2).sqrt()
            noise = (step_noise.pop(0).to(img.value) if step_noise is not None else torch.randn_like(img.value)) * self.sigma
            img = TensorMask(x0.value * a_next.sqrt() + c * pred_noise.value + sigma * noise, x0.mask).apply_mask()
        return img

    @torch.no_grad()
    def sample(self, start: TensorMask, cond: TensorMask, step_noise=None, **kwargs) -> TensorMask:
        if self.is_ddim_sampling:
            return self.ddim_sample(start, cond, step_noise=step_noise, **kwargs)
        assert step_noise is None, "noise injection is implemented for the DDIM sampler"
        return self.p_sample_loop(start, cond, **kwargs)
