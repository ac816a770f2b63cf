"""Conditional affine coupling flow (reference ``modules/flow/layers.py:15-99,199-245``).

Module-level ``forward`` / ``reverse`` compose ``ops.linear`` (FiLM GEMM) with small torch ops and keep
the reference's per-layer API.  ``LVTR`` does not go through them: it hands the stacked parameters
(``CouplingStack.stacked_parameters``) to the fused kernels ``ops.latent_back`` /
``ops.latent_prior_sample`` which run all layers, log_p and the KL in one launch.
Only ``LinearCoupling`` is built — ConvCoupling / spline couplings are not selected by any shipped
config.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...hparams.hp import Hparams
from ...utils.tensormask import TensorMask
from ..activations import get_activation
from ..linear.layers import FiLM
from ..norm import get_norm_fn
from .utils import TensorLogdet


class LinearCoupling(nn.Module):
    def __init__(self, dim: int, flip: bool, hp: Hparams, condition_dim: Optional[int] = None):
        super().__init__()
        hp.check_arg_in_hparams("hidden_dim", "activation", "mean_only", "norm")
        self.dim = dim
        self.mean_only = hp.mean_only
        self.condition_dim = condition_dim
        if condition_dim is not None:
            self.film = FiLM(hp.hidden_dim, in_dim=condition_dim)
        self.linear1 = nn.Linear(dim // 2, hp.hidden_dim, bias=hp.get("bias", True))
        self.linear2 = nn.Linear(hp.hidden_dim, dim // 2 if hp.mean_only else dim, bias=hp.get("bias", True))
        self.norm = get_norm_fn(hp.hidden_dim, hp.norm)
        self.activation = get_activation(hp.activation)
        self.flip = flip
        self.scale_range = hp.get("scale_range", None)
        self.detach_coupling = hp.get("detach_coupling", False)

    def _shift_and_logscale(self, x0: torch.Tensor, c):
        stats = self.norm(F.linear(x0.float(), self.linear1.weight, self.linear1.bias))
        if c is not None and self.condition_dim is not None:
            stats = self.film(stats, c)
        stats = F.linear(self.activation(stats), self.linear2.weight, self.linear2.bias)
        if self.mean_only:
            return stats, torch.zeros_like(stats)
        m, logs = stats.chunk(2, -1)
        if self.scale_range is not None:
            hi, lo = self.scale_range          # unpacked (_max, _min) exactly like the reference (:62-65)
            logs = torch.log(torch.sigmoid(logs) * (hi - lo) + lo)
        return m, logs

    def forward(self, x: TensorLogdet, c: Optional[TensorMask] = None) -> TensorLogdet:
        x0, x1 = x.tensor.value.chunk(2, -1)
        if self.flip:
            x0, x1 = x1, x0
        m, logs = self._shift_and_logscale(x0.detach() if self.detach_coupling else x0, c)
        out = torch.cat([x0, m + x1 * torch.exp(logs)], -1)
        logdet = x.logdet + TensorMask.use_mask(logs, x.tensor.mask)
        return TensorLogdet(TensorMask(out, x.tensor.mask, axis=x.tensor.axis), logdet)

    def reverse(self, x: TensorMask, c: Optional[TensorMask] = None) -> TensorMask:
        x0, x1 = x.value.chunk(2, -1)
        m, logs = self._shift_and_logscale(x0, c)
        x1 = (x1 - m) * torch.exp(-logs)
        if self.flip:
            x0, x1 = x1, x0
        return TensorMask(torch.cat([x0, x1], -1), x.mask, axis=x.axis)


class CouplingStack(nn.Module):
    def __init__(self, dim: int, hp: Hparams, condition_dim: Optional[int] = None) -> None:
        super().__init__()
        hp.check_arg_in_hparams("num_layers", "layer")
        assert hp.num_layers % 2 == 0
        self.identifier = hp.get("identifier", "LinearCoupling")
        if self.identifier != "LinearCoupling":
            raise NotImplementedError(f"{self.identifier}: only LinearCoupling is on the VAE-GSLM hot path")
        self.condition_dim = condition_dim
        self.dim = dim
        # every layer is built with flip=True, as in the reference (:219)
        self.layers = nn.ModuleList([LinearCoupling(dim, True, hp.layer, condition_dim=condition_dim)
                                     for _ in range(hp.num_layers)])

    def forward(self, x: TensorLogdet, c: Optional[TensorMask] = None) -> TensorLogdet:
        for layer in self.layers:
            x = layer(x, c=c)
        return x

    def reverse(self, x: TensorMask, c: Optional[TensorMask] = None) -> TensorMask:
        for layer in reversed(self.layers):
            x = layer.reverse(x, c=c)
        return x

    # ------------------------------------------------------------------ fused-kernel interface
    @property
    def fusable(self) -> bool:
        """the configuration the fused latent kernels implement: conditional, LayerNorm, GELU, σ-range"""
        l0 = self.layers[0]
        return (self.condition_dim is not None and not l0.mean_only and l0.scale_range is not None
                and isinstance(l0.norm, nn.LayerNorm) and isinstance(l0.activation, nn.GELU)
                and not l0.detach_coupling and self.dim == 4 and l0.linear1.out_features == 64
                and len(self.layers) <= 4 and l0.linear1.bias is not None and l0.film.linear.bias is not None)

    def stacked_parameters(self):
        """(w1 [n,Hd,2], b1 [n,Hd], ln_w, ln_b [n,Hd], w2 [n,4,Hd], b2 [n,4]) — differentiable stacks."""
        ls = self.layers
        return (torch.stack([l.linear1.weight for l in ls]), torch.stack([l.linear1.bias for l in ls]),
                torch.stack([l.norm.weight for l in ls]), torch.stack([l.norm.bias for l in ls]),
                torch.stack([l.linear2.weight for l in ls]), torch.stack([l.linear2.bias for l in ls]))

    def film_weight_bias(self):
        """FiLM projections of all layers concatenated along the output dim: [n·2·Hd, cond], [n·2·Hd]."""
        return (torch.cat([l.film.linear.weight for l in self.layers], 0),
                torch.cat([l.film.linear.bias for l in self.layers], 0))

    @property
    def ln_eps(self) -> float:
        return self.layers[0].norm.eps

    @property
    def scale_lo_hi(self):
        hi, lo = self.layers[0].scale_range        # reference unpacks [0.5, 2.0] as (_max, _min)
        return hi, lo
