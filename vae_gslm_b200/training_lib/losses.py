"""Masked losses of the hot path (reference ``training_lib/losses.py:9-73``): sums over valid frames."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .. import ops
from ..utils.tensormask import TensorMask


def masked_loss(x: TensorMask, y: TensorMask, fn: Callable, time_reduction: bool = False,
                batch_reduction: bool = False, batch_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Σ_t mean_c fn(x, y) per sequence over valid frames, then the requested reduction (default: sum)."""
    a = x.flatten().apply_mask().value
    b = y.flatten().apply_mask().value
    out = fn(a, b).mean(-1).sum(-1)
    if batch_weight is not None:
        out = out * batch_weight
    if time_reduction and batch_reduction:
        return out.sum() / x.length.sum()
    if time_reduction:
        return (out / x.length).mean()
    if batch_reduction:
        return out.mean()
    return out.sum()


def masked_ce_loss(x: TensorMask, y: TensorMask, reduction: str = "sum") -> torch.Tensor:
    """token cross-entropy over valid frames (fused softmax-CE kernel; padded frames are ignored)."""
    if reduction != "sum":
        raise NotImplementedError("the hot path uses reduction='sum' (losses.py:34-41)")
    return ops.softmax_ce(x.value, y.value, x.mask)


def l1_loss(a, b):
    return torch.abs(a - b)


def l2_loss(a, b):
    return torch.pow(a - b, 2)


def masked_l1_loss(x: TensorMask, y: TensorMask, time_reduction: bool = False, batch_reduction: bool = False,
                   batch_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    if not (time_reduction or batch_reduction) and batch_weight is None and x.value.dim() == 3:
        return ops.masked_l1(x.value, y.value, x.mask)        # fused kernel (the diffusion loss)
    return masked_loss(x, y, l1_loss, time_reduction, batch_reduction, batch_weight)


def masked_l2_loss(x: TensorMask, y: TensorMask, time_reduction: bool = False, batch_reduction: bool = False,
                   batch_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    return masked_loss(x, y, l2_loss, time_reduction, batch_reduction, batch_weight)
