"""Autoregressive prompt-continuation sampler (reference ``trainers/speech/sampler.py:17-72``).

Loop semantics are the reference's: encode the prompt (stochastic posterior sample), one prefill step
over prompt+BOS (keeping only the last position), then `length − 1` single-token cached steps.  The KV
cache is the pre-allocated in-place ``KVCache`` rather than per-step ``torch.cat``.
"""
from __future__ import annotations

from typing import Mapping, Optional, Tuple

import torch
import torch.nn as nn

from ...utils.tensormask import TensorMask


class ARTRSampler(object):
    def __init__(self, model: nn.Module):
        self.model = model
        self.has_utterance = getattr(model, "utterance_encoder", None) is not None
        self.model_use_tokens = bool(getattr(model, "use_tokens", False))

    @torch.no_grad()
    def __call__(self, length: int, prior: torch.Tensor, temperature: float = 1.0, token_temperature: float = 1.0,
                 truncated_norm: Optional[Tuple[float, float]] = None, return_attn: bool = False,
                 encoder_temperature: float = 1.0, decode: bool = True, greedy: bool = False) -> Mapping:
        model = self.model
        u_c = model.encode_utterance(TensorMask(prior)) if self.has_utterance else None
        prior = model.encode(TensorMask(prior), temperature=encoder_temperature).value
        if not self.model_use_tokens:
            raise NotImplementedError("token-less LVTR variants are outside the VAE-GSLM hot path")
        model.transformer[0].cache_len_hint = prior.shape[1] + 1 + length
        it = {"output": prior, "kv": None}
        frames = [prior]
        for i in range(length):
            it = model.step(it["output"], temperature=temperature, token_temperature=token_temperature,
                            truncated_norm=truncated_norm, past_kv=it["kv"], return_attn=return_attn,
                            push_init_state=(i == 0), greedy=greedy)
            if i == 0:
                it["output"] = it["output"][:, -1:]
            frames.append(it["output"])
        seq = torch.cat(frames, 1)
        outputs = {"frames": seq}
        if decode:
            outputs["output"] = model.decode(TensorMask(seq), u_c=u_c)
        return outputs
