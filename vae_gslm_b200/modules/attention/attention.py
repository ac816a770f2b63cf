"""Self-attention (reference ``modules/attention/attention.py:21-98``) on the fused kernels.

What changes against the reference (all results identical):
  * no dense mask: key padding (per-sequence length), the causal triangle and the ALiBi bias
    −slope_h·(i−j) are evaluated inside the attention kernel;
  * q/k/v are consumed straight from the packed ``in_proj`` output (no chunk / head transposes);
  * with a ``KVCache`` the new k/v are written in place and a single-query decode kernel attends.
``CrossAttention`` (TTS only) is outside the hot path.
"""
from __future__ import annotations

import math
from typing import Any, Mapping, Optional, Tuple

import torch
import torch.nn as nn

from ... import ops
from ...hparams.hp import Hparams
from ...utils.tensormask import TensorMask
from .kvcache import KVCache, LayerKV


class ALiBiBias:
    """Stand-in for the reference's dense ``rpe_bias`` tensor: just the per-head slopes."""

    def __init__(self, slopes: torch.Tensor) -> None:
        self.slopes = slopes


class SelfAttention(nn.Module):
    def __init__(self, dim: int, hp: Hparams) -> None:
        super().__init__()
        hp.check_arg_in_hparams("nheads", "causal")
        self.hp = hp
        self.nheads = hp.nheads
        self.dim = dim
        assert dim % self.nheads == 0
        self.head_dim = dim // self.nheads
        use_bias = bool(hp.get("bias", None))
        self.in_proj = nn.Linear(dim, dim * 3, bias=use_bias)
        self.out_proj = nn.Linear(dim, dim, bias=use_bias)
        self.dropout_p = hp.get("dropout", 0.0)
        if self.dropout_p:
            raise NotImplementedError("attention dropout is not part of the VAE-GSLM configuration")
        if not hp.causal:
            raise NotImplementedError("only causal self-attention is on the VAE-GSLM hot path")

    def _slopes(self, rpe_pair, rpe_bias, outputs) -> Optional[torch.Tensor]:
        if rpe_pair is not None and rpe_pair[0] is not None:
            rpe_id, rpe = rpe_pair
            if rpe_id != "ALiBi":
                raise NotImplementedError(f"positional scheme {rpe_id} is outside the hot path (ALiBi only)")
            outputs["rpe_bias"] = ALiBiBias(rpe.slopes)
            return rpe.slopes
        if rpe_bias is not None:
            if not isinstance(rpe_bias, ALiBiBias):
                raise NotImplementedError("dense rpe_bias tensors are not supported; pass the ALiBi module")
            return rpe_bias.slopes
        return None

    def forward(self, x: TensorMask,
                rpe_pair: Optional[Tuple[str, Any]] = None,
                rpe_bias: Optional[Any] = None,
                return_attn: bool = False,
                past_kv: Optional[Any] = None,
                return_kv: bool = False,
                _residual: Optional[torch.Tensor] = None) -> Mapping[str, Any]:
        outputs = dict()
        slopes = self._slopes(rpe_pair, rpe_bias, outputs)
        B, Tq, _ = x.value.shape
        qkv = ops.linear(x.value, self.in_proj.weight, self.in_proj.bias)          # [B,Tq,3C]
        C = self.dim
        scale = 1.0 / math.sqrt(self.head_dim)
        kv_out = None
        if past_kv is None and not return_kv:
            # training / scoring: full-sequence causal attention with key padding
            o = ops.attention(qkv, self.nheads, x.lengths_i32(), slopes, scale)
        else:
            with torch.no_grad():
                o, kv_out = self._cached(qkv, past_kv, slopes, scale)
        if return_attn:
            outputs["attn"] = self._debug_weights(qkv, past_kv, kv_out, slopes, x)
        # out_proj; padded query rows of `o` are zero.  x + mask(out_proj(o)) when fused with the residual.
        y = ops.linear(o, self.out_proj.weight, self.out_proj.bias, residual=_residual, row_mask=x.mask,
                       mask_before_residual=True)
        outputs["output"] = TensorMask(y, x.mask)
        if return_kv:
            outputs["kv"] = kv_out
        return outputs

    # ------------------------------------------------------------------ KV-cache paths (no grad)
    def _cached(self, qkv: torch.Tensor, past_kv, slopes, scale):
        B, Tq, C3 = qkv.shape
        C = C3 // 3
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        if isinstance(past_kv, Mapping):
            # reference-style dict: concatenate like the reference does (compatibility path)
            k_all = torch.cat([past_kv["key"], k], 1).contiguous()
            v_all = torch.cat([past_kv["value"], v], 1).contiguous()
            o = ops.attention_cached(q.contiguous(), k_all, v_all, self.nheads, k_all.shape[1] - Tq, slopes, scale)
            return o, {"key": k_all.detach(), "value": v_all.detach()}
        if past_kv is None:
            cache = KVCache(1, B, self.nheads, self.head_dim, max(Tq + 512, 1024), qkv.dtype, qkv.device)
            handle, owns = LayerKV(cache, 0), True
        else:
            handle, owns = past_kv, False
        cache, li = handle.cache, handle.index
        pos = cache.length
        assert pos + Tq <= cache.max_len, "KV cache too small (TransformerLayerStack.run grows it before the loop)"
        if Tq == 1:
            o = ops.attention_decode(qkv.view(B, C3), cache.k(li), cache.v(li), pos, slopes, cache.pos_dev, scale,
                                     tickets=cache.tickets)
            o = o.view(B, 1, C)
        else:
            ops.kv_append(k, v, cache.k(li), cache.v(li), pos)
            if pos == 0:
                o = ops.attention_cached(q, k, v, self.nheads, 0, slopes, scale)
            else:
                o = ops.attention_cached(q, cache.k(li), cache.v(li), self.nheads, pos, slopes, scale,
                                         head_major=True, tk=pos + Tq)
        if owns:
            cache.length = pos + Tq
        return o, handle

    @torch.no_grad()
    def _debug_weights(self, qkv, past_kv, kv_out, slopes, x: TensorMask) -> torch.Tensor:
        """softmax attention weights [B,H,Tq,Tk] in fp32 (debug only, as in the reference)."""
        B, Tq, C3 = qkv.shape
        C = C3 // 3
        if kv_out is None:
            k = qkv[..., C:2 * C]
        elif isinstance(kv_out, Mapping):
            k = kv_out["key"]
        else:
            raise NotImplementedError("return_attn is a debug path; use it without the in-place KV cache")
        Tk = k.shape[1]
        H, D = self.nheads, self.head_dim
        qh = qkv[..., :C].float().view(B, Tq, H, D).transpose(1, 2)
        kh = k.float().view(B, Tk, H, D).transpose(1, 2)
        s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
        i = torch.arange(Tk - Tq, Tk, device=qkv.device)[:, None]
        j = torch.arange(Tk, device=qkv.device)[None, :]
        ok = (j <= i)[None, None]
        if Tk == Tq:
            ok = ok & x.mask[:, None, None, :]
        if slopes is not None:
            s = s - slopes.view(1, H, 1, 1) * (i - j).clamp(min=0)[None, None]
        return torch.softmax(s.masked_fill(~ok, float("-inf")), -1)

    def custom_weight_init(self, init_std: float):
        std = init_std / math.sqrt(self.dim / 3)
        self.in_proj.weight.data.uniform_(-std, std)
        self.out_proj.weight.data.uniform_(-std, std)


class CrossAttention(nn.Module):
    def __init__(self, *args, **kwargs) -> None:
        super().__init__()
        raise NotImplementedError("CrossAttention belongs to the TTS variant, outside the VAE-GSLM hot path")
