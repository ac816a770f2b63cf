"""Pre-allocated KV cache for the cached ``LVTR.step`` loop.

The reference keeps, per layer, ``{"key": [B,Tk,C], "value": [B,Tk,C]}`` and re-``torch.cat``s the whole
thing every step (attention.py:56-58,81-85).  Here all layers share one head-major allocation
``[L, 2, B, H, Tmax, D]`` that is written in place by the attention kernels; ``LayerKV`` is the
per-layer handle that travels through the reference's ``past_kv`` / ``kv`` plumbing and can
materialise the reference layout on demand (``handle["key"]``).
"""
from __future__ import annotations

from typing import List, Optional

import torch


class KVCache:
    def __init__(self, n_layers: int, batch: int, nheads: int, head_dim: int, max_len: int,
                 dtype: torch.dtype, device) -> None:
        self.n_layers, self.batch, self.nheads, self.head_dim = n_layers, batch, nheads, head_dim
        self.max_len = int(max_len)
        self.buf = torch.empty((n_layers, 2, batch, nheads, self.max_len, head_dim), dtype=dtype, device=device)
        self.length = 0                       # host-side number of valid positions (same for all layers)
        self.pos_dev: Optional[torch.Tensor] = None   # optional device-resident copy for CUDA-graph replay
        # split-KV decode: per-(sequence, head) arrival counters, zero between launches (vg_attn_decode merges in-kernel)
        self.tickets = torch.zeros(batch * nheads, dtype=torch.int32, device=device)

    def k(self, layer: int) -> torch.Tensor:
        return self.buf[layer, 0]

    def v(self, layer: int) -> torch.Tensor:
        return self.buf[layer, 1]

    def layers(self) -> List["LayerKV"]:
        return [LayerKV(self, i) for i in range(self.n_layers)]

    def ensure(self, needed: int) -> None:
        """grow (amortised doubling) so that ``needed`` positions fit; keeps the valid prefix."""
        if needed <= self.max_len:
            return
        new_len = max(needed, 2 * self.max_len)
        buf = torch.empty((self.n_layers, 2, self.batch, self.nheads, new_len, self.head_dim),
                          dtype=self.buf.dtype, device=self.buf.device)
        buf[:, :, :, :, : self.length] = self.buf[:, :, :, :, : self.length]
        self.buf, self.max_len = buf, new_len

    @classmethod
    def from_reference(cls, past_kv, nheads: int, extra: int = 0) -> "KVCache":
        """import a reference-style list of {"key","value"} [B,Tk,C] dicts."""
        key0 = past_kv[0]["key"]
        B, Tk, Cdim = key0.shape
        D = Cdim // nheads
        cache = cls(len(past_kv), B, nheads, D, Tk + extra, key0.dtype, key0.device)
        for i, kv in enumerate(past_kv):
            cache.buf[i, 0, :, :, :Tk] = kv["key"].view(B, Tk, nheads, D).transpose(1, 2)
            cache.buf[i, 1, :, :, :Tk] = kv["value"].view(B, Tk, nheads, D).transpose(1, 2)
        cache.length = Tk
        return cache


class LayerKV:
    """One layer's window into a KVCache; quacks like the reference's per-layer kv dict."""

    def __init__(self, cache: KVCache, index: int) -> None:
        self.cache, self.index = cache, index

    def _ref_layout(self, which: int) -> torch.Tensor:
        c = self.cache
        t = c.buf[self.index, which, :, :, : c.length]              # [B,H,Tk,D]
        return t.transpose(1, 2).reshape(c.batch, c.length, c.nheads * c.head_dim)

    def __getitem__(self, name: str) -> torch.Tensor:
        if name == "key":
            return self._ref_layout(0)
        if name == "value":
            return self._ref_layout(1)
        raise KeyError(name)

    def keys(self):
        return ("key", "value")
