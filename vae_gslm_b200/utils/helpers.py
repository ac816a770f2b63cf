"""The two helpers of the reference's ``utils/helpers.py`` that sit on the hot path (:138-150)."""
from __future__ import annotations

import torch


def get_padding(kernel_size: int, dilation: int = 1, stride: int = 1, causal: bool = False, future: bool = False):
    """'same' padding; (2p, 0) for causal convolutions, (0, 2p) for future-looking ones."""
    p = int(((kernel_size - 1) * dilation + 1 - stride) / 2)
    if causal:
        return (2 * p, 0)
    if future:
        return (0, 2 * p)
    return p


def make_padding_mask(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[B,Ta] x [B,Tb] → [B,Ta,Tb]: key validity broadcast over queries (query padding is NOT masked)."""
    return b.unsqueeze(-2).expand(-1, a.size(1), -1)
