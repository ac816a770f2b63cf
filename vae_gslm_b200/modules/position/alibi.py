"""ALiBi (reference ``modules/position/alibi.py:6-33``).

The reference precomputes a dense [H, maxpos, maxpos] buffer of −slope_h·|i−j| (64 MB at the repo
config) and adds it to a dense [B,H,T,T] mask in every layer.  Here only the H slopes are kept: the
attention kernels compute −slope_h·(i−j) in registers (only j ≤ i survives the causal mask), so there
is no ``maxpos`` limit and no O(T²) memory.  ``forward`` still materialises the reference buffer slice
on demand for debugging / API parity.
"""
import torch
import torch.nn as nn

from ...ops import alibi_slopes


class ALiBi(nn.Module):
    def __init__(self, nheads: int, maxpos: int = 10000) -> None:
        super().__init__()
        self.nheads = nheads
        self.maxpos = maxpos
        self.register_buffer("slopes", torch.tensor(alibi_slopes(nheads), dtype=torch.float32), persistent=False)

    def get_slopes(self, n: int):
        return alibi_slopes(n)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[H, Tq, Tk] additive bias for a [B,H,Tq,Tk]-shaped argument (debug path; not used by the kernels)."""
        i = torch.arange(x.size(2), device=self.slopes.device)[:, None]
        j = torch.arange(x.size(3), device=self.slopes.device)[None, :]
        return -self.slopes[:, None, None] * (j - i).abs()[None]
