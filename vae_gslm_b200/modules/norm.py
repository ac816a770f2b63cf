"""Normalisation layers (reference ``modules/norm.py``).

``RMSNorm`` runs the fused CUDA kernel (fp32 statistics, parameter named ``scale`` as in the
checkpoint contract).  ``InstanceNorm`` — which in the reference is a per-time-step LayerNorm over the
channel dimension of a B,C,T tensor (:35-47) — holds the parameters of the conv blocks' channel norm; in the
residual blocks of the conv encoder / UNet it is evaluated inside the fused ``dwconv_ln`` kernel
(``modules/conv/layers.py``), and as a torch module only by the small utterance encoder.
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..hparams.hp import Hparams


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5) -> None:
        super().__init__()
        self.eps = eps
        self.scale = nn.Parameter(torch.ones(dim))

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """y = scale · x · rsqrt(mean(x²) + eps); rows where ``mask`` is False are written as zeros."""
        return ops.rmsnorm(x, self.scale, self.eps, mask, out_dtype)


class InstanceNorm(nn.Module):
    """Channel LayerNorm of a B,C,T tensor, statistics per (b, t); unbiased variance like the reference."""

    def __init__(self, dim: int, eps: float = 1e-5) -> None:
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.float()
        var, mean = torch.var_mean(x, dim=1, keepdim=True)
        return self.weight[:, None] * ((x - mean) * torch.rsqrt(var + self.eps)) + self.bias[:, None]


def get_norm_fn(dim: int, hp: Hparams) -> nn.Module:
    kind = hp.identifier
    if kind == "RMSNorm":
        return RMSNorm(dim, eps=hp.eps)
    if kind == "LayerNorm":
        return nn.LayerNorm(dim, eps=hp.eps)
    if kind == "InstanceNorm":
        return InstanceNorm(dim, eps=hp.eps)
    if kind == "GroupNorm":
        return nn.GroupNorm(hp.num_groups, dim, eps=hp.eps)
    if kind == "Identity":
        return nn.Identity()
    raise ValueError(f"{kind} not in the usable normalization function lists.")
