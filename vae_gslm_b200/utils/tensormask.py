"""(value, validity-mask) carrier used by every hot-path signature.

API mirror of the reference's ``utils/tensormask.py:7-229``: ``mask[b, t]`` is True for real frames,
padding sits at the right; ``axis`` says whether time is dimension 1 (B,T,...) or 2 (B,C,T).
The kernels consume the mask as a flat uint8 row mask and as per-sequence int32 lengths
(``row_mask_u8`` / ``lengths_i32`` below are additions for that purpose).
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch


_PASS = [0]


def new_pass() -> None:
    """called at the start of every transformer-stack pass: lengths memoised on a mask tensor (``lengths_i32``) are valid
    for ONE pass only.  Static input buffers are refilled in place between steps (TrainStep.load), and a CUDA-graph capture
    must contain the reduction itself, not a tensor memoised during warm-up."""
    _PASS[0] += 1


class TensorMask(object):
    def __init__(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, axis: int = 1) -> None:
        if axis not in (1, 2):
            raise AssertionError("Only Support B T ..., B C T")
        if mask is None:
            mask = torch.ones(x.shape[0], x.shape[1], dtype=torch.bool, device=x.device)
        if mask.dim() != 2:
            raise AssertionError("mask must be [B, T]")
        t_dim = 1 if axis == 1 else 2
        if (x.shape[0], x.shape[t_dim]) != tuple(mask.shape):
            raise AssertionError(f"value {tuple(x.shape)} and mask {tuple(mask.shape)} disagree (axis={axis})")
        self.value = x
        self.mask = mask
        self.axis = axis

    # ------------------------------------------------------------------ construction
    @classmethod
    def fromlength(cls, x: torch.Tensor, length: torch.Tensor, axis: int = 1) -> "TensorMask":
        steps = torch.arange(x.shape[axis], device=x.device)
        return cls(x, steps.unsqueeze(0) < length.unsqueeze(1), axis)

    @classmethod
    def use_mask(cls, x: torch.Tensor, mask: torch.Tensor, mask_value: float = 0) -> torch.Tensor:
        return cls(x, mask).apply_mask(mask_value).value

    @classmethod
    def resize_length(cls, length: torch.Tensor, ratio) -> torch.Tensor:
        return torch.ceil(length.float() * ratio).long()

    # ------------------------------------------------------------------ kernel-facing views
    def row_mask_u8(self) -> torch.Tensor:
        return self.mask.contiguous().view(torch.uint8).reshape(-1)

    def lengths_i32(self) -> torch.Tensor:
        # memoised on the mask tensor for the current pass (16 layers share one mask): one reduction per pass
        key = (_PASS[0], self.mask._version)
        cached = getattr(self.mask, "_vg_lengths", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        lengths = self.mask.sum(-1, dtype=torch.int32)
        try:
            self.mask._vg_lengths = (key, lengths)
        except Exception:
            pass
        return lengths

    # ------------------------------------------------------------------ masking
    def apply_mask(self, mask_value: float = 0) -> "TensorMask":
        assert self.axis == 1
        m = self.mask.reshape(self.mask.shape + (1,) * (self.value.dim() - 2))
        return TensorMask(torch.where(m, self.value, mask_value), self.mask)

    # ------------------------------------------------------------------ shape helpers
    def __len__(self):
        return len(self.value)

    def __repr__(self):
        return repr({"value": self.value, "mask": self.mask, "axis": self.axis})

    def size(self, i: Optional[int] = None):
        return self.value.size() if i is None else self.value.size(i)

    @property
    def length(self) -> torch.Tensor:
        return self.mask.long().sum(-1)

    @property
    def device(self):
        return self.value.device

    def flatten(self) -> "TensorMask":
        assert self.axis == 1
        b, t = self.value.shape[:2]
        return TensorMask(self.value.reshape(b, t, -1), self.mask)

    def transpose(self, a: int = -1, b: int = -2) -> "TensorMask":
        return TensorMask(self.value.transpose(a, b), self.mask, axis=3 - self.axis)

    def squeeze(self, dim: Optional[int] = None) -> "TensorMask":
        return TensorMask(self.value.squeeze() if dim is None else self.value.squeeze(dim), self.mask)

    def expand(self) -> "TensorMask":
        return TensorMask(self.value.unsqueeze(-1), self.mask)

    def long(self) -> "TensorMask":
        return TensorMask(self.value.long(), self.mask)

    def abs(self) -> "TensorMask":
        return TensorMask(self.value.abs(), self.mask)

    def split(self, n: int) -> Tuple["TensorMask", "TensorMask"]:
        return TensorMask(self.value[..., :n], self.mask), TensorMask(self.value[..., n:], self.mask)

    def tolist(self, detach: bool = True) -> List[torch.Tensor]:
        assert self.axis == 1
        rows = [v[m] for v, m in zip(self.value, self.mask)]
        return [r.detach() for r in rows] if detach else rows

    # ------------------------------------------------------------------ device / autograd plumbing
    def cuda(self) -> "TensorMask":
        return TensorMask(self.value.cuda(), self.mask.cuda(), self.axis)

    def to(self, device, non_blocking: bool = False) -> "TensorMask":
        return TensorMask(self.value.to(device, non_blocking=non_blocking),
                          self.mask.to(device, non_blocking=non_blocking), self.axis)

    def detach(self) -> "TensorMask":
        return TensorMask(self.value.detach(), self.mask.detach(), self.axis)

    # ------------------------------------------------------------------ time-axis edits
    def push(self, tm: Union[torch.Tensor, "TensorMask"]) -> "TensorMask":
        """prepend frames (all valid when a bare tensor is given)."""
        assert self.axis == 1
        tm = tm if isinstance(tm, TensorMask) else TensorMask(tm)
        return TensorMask(torch.cat([tm.value, self.value], 1), torch.cat([tm.mask, self.mask], 1))

    def append(self, tm: Union[torch.Tensor, "TensorMask"]) -> "TensorMask":
        assert self.axis == 1
        tm = tm if isinstance(tm, TensorMask) else TensorMask(tm)
        return TensorMask(torch.cat([self.value, tm.value], 1), torch.cat([self.mask, tm.mask], 1))

    def pop(self, n: Union[int, torch.Tensor] = 1) -> "TensorMask":
        """drop the last n frames; every sequence gets n shorter."""
        assert self.axis == 1
        return TensorMask.fromlength(self.value[:, :-n], self.length - n)

    def pop_left(self, n: Union[int, torch.Tensor] = 1) -> "TensorMask":
        return TensorMask.fromlength(self.value[:, n:], self.length - n)

    def cat(self, other: Union[torch.Tensor, "TensorMask"]) -> "TensorMask":
        """channel concat (dim 2 for B,T,C; dim 1 for B,C,T)."""
        other = other.value if isinstance(other, TensorMask) else other
        return TensorMask(torch.cat([self.value, other], 3 - self.axis), self.mask, axis=self.axis)

    # ------------------------------------------------------------------ reductions / arithmetic
    def mean(self) -> torch.Tensor:
        """mean over valid frames and channels."""
        assert self.axis == 1
        x = self.flatten().apply_mask().value
        return (x / x.size(-1)).sum() / self.length.sum()

    def _binary(self, other, fn) -> "TensorMask":
        other = other.value if isinstance(other, TensorMask) else other
        return TensorMask(fn(self.value, other), self.mask, axis=self.axis)

    def __truediv__(self, other):
        return self._binary(other, torch.div)

    def __mul__(self, other):
        return self._binary(other, torch.mul)

    def __add__(self, other):
        return self._binary(other, torch.add)

    def __sub__(self, other):
        return self._binary(other, torch.sub)

    def batch_time_shuffle(self) -> "TensorMask":
        """randomly permute the valid frames across batch and time (padding stays put)."""
        assert self.axis == 1 and self.value.dim() == 3
        b, t, c = self.value.shape
        slots = torch.arange(b * t, device=self.device).reshape(b, t)[self.mask]
        slots = slots[torch.randperm(len(slots), device=self.device)]
        out = torch.zeros(b * t, c, dtype=self.value.dtype, device=self.device)
        out[slots] = self.value[self.mask]
        return TensorMask(out.reshape(b, t, c), self.mask).apply_mask()
