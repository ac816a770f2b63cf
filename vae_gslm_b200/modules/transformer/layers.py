"""Transformer layer and stack (reference ``modules/transformer/layers.py:13-195``) on the fused kernels.

Per pre-LN layer the launch sequence is: RMSNorm(+mask) → QKV GEMM → attention → out-proj GEMM with
residual+mask epilogue → RMSNorm → FFN1 GEMM (bias+GELU epilogue, pre-activation saved) → FFN2 GEMM
(bias+residual+mask epilogue): 7 kernels against the reference's ~40.
"""
from __future__ import annotations

import math
from typing import Any, List, Mapping, Optional, Tuple

import torch
import torch.nn as nn

from ... import ops
from ...hparams.hp import Hparams
from ...utils.tensormask import TensorMask
from ..activations import get_activation
from ..attention.attention import SelfAttention
from ..attention.kvcache import KVCache, LayerKV
from ..norm import RMSNorm, get_norm_fn
from ..position.embedding import get_positional_encoding


def _norm_with_residual(norm: nn.Module, x: torch.Tensor, mask: Optional[torch.Tensor]):
    """(residual input, norm(x)) of a pre-LN block.  For RMSNorm the residual is an autograd alias of x whose
    gradient is added inside the fused RMSNorm backward kernel (ops.rmsnorm_residual) instead of a separate add."""
    if isinstance(norm, RMSNorm) and torch.is_grad_enabled() and x.requires_grad:
        return ops.rmsnorm_residual(x, norm.scale, norm.eps, mask)
    return x, _apply_norm(norm, x, mask)


def _apply_norm(norm: nn.Module, x: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    if isinstance(norm, RMSNorm):
        return norm(x, mask)
    y = norm(x).to(x.dtype)
    return ops.mask_rows(y, mask) if mask is not None else y


class TransformerLayer(nn.Module):
    def __init__(self, hp: Hparams) -> None:
        super().__init__()
        hp.check_arg_in_hparams("ffd_size", "norm", "activation", "dim", "self_attn")
        if hp.get("dropout", 0.0):
            raise NotImplementedError("dropout is 0 in the VAE-GSLM configuration; not built")
        if hp.has("cross_attn"):
            raise NotImplementedError("cross-attention layers belong to the TTS variant (outside the hot path)")
        self.hp = hp
        self.preln = hp.get("preln", True)
        self.self_attn = SelfAttention(hp.dim, hp.self_attn)
        self.cross_attn = None
        # NB: the FFN bias defaults to True — the YAML's `bias: false` sits one level up (SURVEY §8a note B)
        self.linear1 = nn.Linear(hp.dim, hp.ffd_size, bias=hp.get("bias", True))
        self.linear2 = nn.Linear(hp.ffd_size, hp.dim, bias=hp.get("bias", True))
        self.norm1 = get_norm_fn(hp.dim, hp.norm)
        self.norm3 = get_norm_fn(hp.dim, hp.norm)
        self.activation = get_activation(hp.activation)
        self._act_id = ops.ACT_IDS.get(hp.activation.identifier)

    def _ffn(self, n: torch.Tensor, residual: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if self._act_id is not None:
            return ops.ffn(n, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                           residual=residual, row_mask=mask, act=self._act_id)
        h = self.activation(ops.linear(n, self.linear1.weight, self.linear1.bias))
        return ops.linear(h, self.linear2.weight, self.linear2.bias, residual=residual, row_mask=mask)

    def forward(self, tgt: TensorMask, memory: Optional[TensorMask] = None,
                rpe_pair: Optional[Tuple[str, Any]] = None, rpe_bias: Optional[Any] = None,
                past_kv: Optional[Any] = None, return_attn: bool = False, return_kv: bool = False) -> Mapping:
        output = dict()
        x, mask = tgt.value, tgt.mask
        if self.preln:
            x_res, n1 = _norm_with_residual(self.norm1, x, mask)
            sa = self.self_attn(TensorMask(n1, mask), past_kv=past_kv, rpe_pair=rpe_pair, rpe_bias=rpe_bias,
                                return_attn=return_attn, return_kv=return_kv, _residual=x_res)
            h1_res, n3 = _norm_with_residual(self.norm3, sa["output"].value, None)   # sa output = x + mask(attn)
            out = self._ffn(n3, h1_res, mask)
        else:
            sa = self.self_attn(tgt, past_kv=past_kv, rpe_pair=rpe_pair, rpe_bias=rpe_bias,
                                return_attn=return_attn, return_kv=return_kv, _residual=x)
            h1 = _apply_norm(self.norm1, sa["output"].value, None)
            h2 = self._ffn(h1, h1, None)
            out = _apply_norm(self.norm3, h2, mask)
        if "rpe_bias" in sa:
            output["rpe_bias"] = sa["rpe_bias"]
        output["output"] = TensorMask(out, mask)
        if return_attn:
            output["self_attn"] = sa["attn"]
        if return_kv:
            output["kv"] = sa["kv"]
        return output


class TransformerLayerStack(nn.Module):
    def __init__(self, hp: Hparams, input_dim: Optional[int] = None, output_dim: Optional[int] = None,
                 memory_dim: Optional[int] = None) -> None:
        super().__init__()
        hp.check_arg_in_hparams("num_layers", "layer")
        self.hp = hp
        self.layers = nn.ModuleList([TransformerLayer(hp.layer) for _ in range(hp.num_layers)])
        use_bias = hp.get("bias", True)
        self.linear = nn.Linear(input_dim, hp.layer.dim, bias=use_bias) if input_dim is not None else None
        self.out = nn.Linear(hp.layer.dim, output_dim, bias=use_bias) if output_dim is not None else None
        self.memory_linear = None
        self.is_cross_attn = False
        self.final_norm = get_norm_fn(hp.layer.dim, hp.layer.norm) if hp.get("final_ln", True) else None
        self.first_norm = get_norm_fn(hp.layer.dim, hp.layer.norm) if hp.get("first_ln", False) else None
        self.rpe, self.rpe_id = None, None
        if hp.get("rpe", False):
            self.rpe_id = hp.rpe.identifier
            self.rpe = get_positional_encoding(self.rpe_id, hp.rpe, hp.layer.dim, hp.layer.self_attn.nheads)
        self.compute_dtype = torch.float32      # activation stream dtype: float32 (parity) or bfloat16
        self.cache_len_hint = 0                 # total positions a fresh KV cache should hold

    def run(self, tgt: TensorMask, memory: Optional[TensorMask] = None, past_kv: Optional[List] = None,
            return_attn: bool = False, return_kv: bool = False) -> Mapping[str, Any]:
        from ...utils.tensormask import new_pass
        new_pass()                                  # sequence lengths memoised on the mask tensor are per-pass
        outputs = {"output": []}
        if return_attn:
            outputs["self_attn"] = []
        mask = tgt.mask
        x = tgt.value.to(self.compute_dtype)
        B, Tq = x.shape[0], x.shape[1]
        # ---- KV cache set-up: one shared allocation, per-layer handles
        cache = None
        if return_kv or past_kv is not None:
            if past_kv is None:
                hps = self.hp.layer
                nh = hps.self_attn.nheads
                cache = KVCache(len(self.layers), B, nh, hps.dim // nh, max(self.cache_len_hint, Tq + 64),
                                self.compute_dtype, x.device)
                past_kv = cache.layers()
            elif isinstance(past_kv[0], LayerKV):
                cache = past_kv[0].cache
            elif isinstance(past_kv[0], Mapping) and return_kv:
                nh = self.hp.layer.self_attn.nheads
                cache = KVCache.from_reference(past_kv, nh, extra=max(Tq + 64, self.cache_len_hint))
                past_kv = cache.layers()
            if cache is not None:
                cache.ensure(cache.length + Tq)
        if past_kv is None:
            past_kv = [None] * len(self.layers)

        if self.linear is not None:
            x = ops.linear(x, self.linear.weight, self.linear.bias, row_mask=mask)
        if self.first_norm is not None:
            x = _apply_norm(self.first_norm, x, mask)
        rpe_pair = (self.rpe_id, self.rpe)
        rpe_bias = None
        output = TensorMask(x, mask)
        output_layers = []
        if return_kv:
            outputs["kv"] = []
        for i, mod in enumerate(self.layers):
            res = mod(output, memory, rpe_pair=rpe_pair, rpe_bias=rpe_bias, past_kv=past_kv[i],
                      return_attn=return_attn, return_kv=return_kv)
            if "rpe_bias" in res:
                rpe_pair, rpe_bias = None, res["rpe_bias"]
            if return_attn:
                outputs["self_attn"].append(res["self_attn"].detach())
            if return_kv:
                outputs["kv"].append(res["kv"])
            output = res["output"]
            output_layers.append(output)
        if cache is not None:
            cache.length += Tq
            if cache.pos_dev is not None:
                ops.L.call("vg_add_i32", ops.L.ptr(cache.pos_dev), Tq, ops.L.stream())
        if self.final_norm is not None:
            # not re-masked: padded rows are zero already and RMSNorm(0) = 0
            output = TensorMask(_apply_norm(self.final_norm, output.value, None), output.mask)
            output_layers.append(output)
        if self.out is not None:
            output = TensorMask(ops.linear(output.value, self.out.weight, self.out.bias, row_mask=mask), mask)
        outputs["output"] = output
        outputs["layers"] = output_layers
        return outputs

    def forward(self, tgt: TensorMask, memory: Optional[TensorMask] = None) -> TensorMask:
        return self.run(tgt, memory=memory)["output"]

    def custom_weight_init(self, init_std: float):
        pass    # only T5RPE has stack-level weights in the reference (layers.py:201-204); not built here
