"""Linear-family modules of the hot path (reference ``modules/linear/layers.py``): GaussianParameterize
(:54-147), Embedding (:150-157), Linear (:184-193), TimeAggregation (:260-262), FiLM (:265-292).

Standing alone, each module issues its GEMM through ``ops.linear`` (libvgslm) and finishes with small
torch elementwise ops; inside ``LVTR.forward`` the posterior/prior heads, the embedding and the fuser
are consumed by the fused latent kernels instead (``ops.latent_front`` / ``ops.latent_back``), reading
these modules' parameters directly.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...utils.attr import AttrDict
from ...utils.tensormask import TensorMask

_ACT_OF_MODULE = {nn.ReLU: ops.ACT_RELU, nn.GELU: ops.ACT_GELU, nn.Identity: ops.ACT_NONE}


class GaussianParameterize(nn.Module):
    """mean / logstd heads + reparameterised sample.  NB: parameterised by ``logstd`` (not logvar):
    sample = mean + exp(logstd)·eps·temperature (linear/layers.py:122-128)."""

    def __init__(self, in_dim: int, dim: int, bias: bool = True, std: Optional[float] = None,
                 std_range: Optional[Tuple[float, float]] = None,
                 truncated_norm: Optional[Tuple[float, float]] = None, total_std: Optional[float] = None,
                 use_tanh: bool = False, use_relu: bool = False, normalization: bool = False,
                 mean: Optional[float] = None):
        super().__init__()
        self._mean = mean
        self.dim = dim
        if mean is None:
            self.mean = nn.Linear(in_dim, dim, bias=bias)
        self.std = std
        self.truncated_norm = truncated_norm
        if std is None:
            self.logstd = nn.Linear(in_dim, dim, bias=bias)
        self.std_range = None
        if std_range is not None:
            assert std is None and len(std_range) == 2
            self.std_range = std_range
        self.total_std = total_std
        if total_std is not None:
            assert std is None and std_range is None
        self.use_tanh, self.use_relu, self.normalization = use_tanh, use_relu, normalization

    @property
    def is_plain(self) -> bool:
        """True for the variant the fused latent kernels implement (learned mean and logstd, no extras)."""
        return (self._mean is None and self.std is None and self.std_range is None and self.total_std is None
                and self.truncated_norm is None and not (self.use_tanh or self.use_relu or self.normalization))

    def forward(self, x: TensorMask, temperature: float = 1.0,
                truncated_norm: Optional[Tuple[float, float]] = None,
                eps: Optional[torch.Tensor] = None) -> AttrDict:
        v = x.value
        if self._mean is None:
            mean = ops.linear(v, self.mean.weight, self.mean.bias, out_dtype=torch.float32)
        else:
            mean = torch.full(v.shape[:2] + (self.dim,), self._mean, device=v.device)
        if self.normalization:
            mean = F.normalize(mean, p=2.0, dim=-1)
        if self.use_relu:
            mean = F.relu(mean)
        if self.use_tanh:
            mean = torch.tanh(mean) * 0.5
        if self.std is None:
            logstd = ops.linear(v, self.logstd.weight, self.logstd.bias, out_dtype=torch.float32)
            if self.std_range is not None:
                hi, lo = self.std_range      # unpacked (_max, _min) exactly like the reference
                logstd = torch.log(torch.sigmoid(logstd) * (hi - lo) + lo)
        else:
            logstd = torch.log(torch.full(mean.size(), self.std, device=v.device))
        noise = torch.randn_like(mean) if eps is None else eps.to(mean.dtype)
        for rng in (self.truncated_norm, truncated_norm):
            if rng is not None:
                nn.init.trunc_normal_(noise, a=rng[0], b=rng[1])
        std = torch.exp(logstd.float())
        if self.total_std is not None:
            std = std / std.sum(-1, keepdim=True) * self.total_std * std.size(-1)
            logstd = torch.log(std)
        sample = mean + noise * std * temperature
        return AttrDict(mean=TensorMask(mean, x.mask), logstd=TensorMask(logstd, x.mask),
                        sample=TensorMask(sample, x.mask))


class Embedding(nn.Embedding):
    def forward(self, x: TensorMask) -> TensorMask:
        return TensorMask(super().forward(x.value), x.mask).apply_mask()

    def custom_weight_init(self, init_std: float):
        self._fill_padding_idx_with_zero()
        self.weight.data.uniform_(-1.0, 1.0)


class Linear(nn.Module):
    """act(x·Wᵀ + b) on a TensorMask, NOT masked (padded rows carry act(b))."""

    def __init__(self, in_dim: int, out_dim: int, bias: bool = True, activation=nn.Identity()) -> None:
        super().__init__()
        self.linear = nn.Linear(in_dim, out_dim, bias=bias)
        self.activation = activation

    def forward(self, x: TensorMask) -> TensorMask:
        act = _ACT_OF_MODULE.get(type(self.activation))
        v = x.value
        if v.shape[-1] < 8:          # K too small for the GEMM kernels' vector paths (token_fuser, 4→64)
            y = self.activation(F.linear(v.float(), self.linear.weight, self.linear.bias))
        elif act is None:
            y = self.activation(ops.linear(v, self.linear.weight, self.linear.bias))
        else:
            y = ops.linear(v, self.linear.weight, self.linear.bias, act=act)
        return TensorMask(y, x.mask)


class TimeAggregation(nn.Module):
    def forward(self, x: TensorMask) -> torch.Tensor:
        return x.flatten().apply_mask().value.sum(1) / x.length[..., None]


class FiLM(nn.Module):
    def __init__(self, dim: int, bias: bool = True, time_first: bool = True, in_dim: Optional[int] = None):
        super().__init__()
        in_dim = dim if in_dim is None else in_dim
        self.linear = nn.Linear(in_dim, dim * 2, bias=bias) if time_first else nn.Conv1d(in_dim, dim * 2, 1, bias=bias)
        self.time_first = time_first

    def forward(self, x: Union[torch.Tensor, TensorMask], c: Union[torch.Tensor, TensorMask]):
        y = x.value if isinstance(x, TensorMask) else x
        c = c.value if isinstance(c, TensorMask) else c
        if self.time_first:
            wb = ops.linear(c, self.linear.weight, self.linear.bias, out_dtype=torch.float32)
            weight, bias = wb.chunk(2, -1)
        else:
            weight, bias = self.linear(c).chunk(2, 1)
        y = weight * y + bias
        if isinstance(x, TensorMask):
            return TensorMask(y, x.mask, axis=1 if self.time_first else 2)
        return y
