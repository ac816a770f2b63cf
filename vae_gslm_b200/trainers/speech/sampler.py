"""Autoregressive prompt-continuation sampler (reference ``trainers/speech/sampler.py:17-72``).

Loop semantics are the reference's: encode the prompt (stochastic posterior sample), one prefill step
over prompt+BOS (keeping only the last position), then `length − 1` single-token cached steps.  What
changes: the KV cache is the pre-allocated in-place ``KVCache`` (no per-step ``torch.cat``) and the
single-token step — ~150 small kernels — is captured once into a CUDA graph and replayed; the cache
position lives in device memory (``KVCache.pos_dev``) so that one graph serves every step.
"""
from __future__ import annotations

from typing import Mapping, Optional, Tuple

import torch
import torch.nn as nn

from ...utils.tensormask import TensorMask


class GraphedStep:
    """One single-token ``LVTR.step`` captured into a CUDA graph.

    ``state`` [B,1,1+L] is a static buffer: every replay consumes it, writes the new (token, z) frame back
    into it and returns that buffer.  Random draws (prior eps, token uniforms) are torch RNG calls inside the
    captured region, which CUDA-graph-safe generators re-seed on every replay."""

    def __init__(self, model: nn.Module, state: torch.Tensor, kv, **step_kwargs) -> None:
        self.model = model
        self.cache = kv[0].cache
        cache = self.cache
        self.state = state.clone()
        if cache.pos_dev is None:
            cache.pos_dev = torch.tensor([cache.length], dtype=torch.int32, device=state.device)
        else:
            cache.pos_dev.fill_(cache.length)
        assert cache.length + 2 <= cache.max_len, "grow the KV cache before capturing (cache_len_hint)"
        self.graph = torch.cuda.CUDAGraph()
        host_len = cache.length
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            out = model.step(self.state, past_kv=kv, **step_kwargs)
            self.state.copy_(out["output"])
        cache.length = host_len            # capture ran the host bookkeeping once but no kernel
        self.kv = kv

    def __call__(self) -> torch.Tensor:
        assert self.cache.length + 1 <= self.cache.max_len, "KV cache exhausted"
        self.graph.replay()
        self.cache.length += 1             # pos_dev was advanced inside the graph (vg_add_i32)
        return self.state


class GraphedDecode:
    """``LVTR.decode`` (DDIM sampling of mel frames, lvtr.py:288-306 → ddpm.py:284-321) captured into ONE CUDA graph.

    The sampling schedule is fixed by the configuration, so the whole loop — ``sampling_timesteps`` UNet passes of ~60
    short kernels each — is unrolled into the graph; time steps and DDIM coefficients are baked in, the noise draws are
    graph-safe torch RNG calls.  ``frames`` [B,T,1+L], ``mask`` [B,T] and ``u_c`` [B,E] are static input buffers."""

    def __init__(self, model: nn.Module, frames: torch.Tensor, mask: torch.Tensor, u_c: Optional[torch.Tensor]) -> None:
        self.model = model
        self.frames, self.mask = frames.clone(), mask.clone()
        self.u_c = u_c.clone() if u_c is not None else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            model.decode(TensorMask(self.frames, self.mask), u_c=self.u_c)          # warm-up (lazy initialisations)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = model.decode(TensorMask(self.frames, self.mask), u_c=self.u_c).value

    def __call__(self, frames: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                 u_c: Optional[torch.Tensor] = None) -> torch.Tensor:
        if frames is not None:
            self.frames.copy_(frames)
        if mask is not None:
            self.mask.copy_(mask)
        if u_c is not None and self.u_c is not None:
            self.u_c.copy_(u_c)
        self.graph.replay()
        return self.out


class ARTRSampler(object):
    def __init__(self, model: nn.Module, use_cuda_graph: bool = True):
        self.model = model
        self.has_utterance = getattr(model, "utterance_encoder", None) is not None
        self.model_use_tokens = bool(getattr(model, "use_tokens", False))
        self.use_cuda_graph = use_cuda_graph

    @torch.no_grad()
    def __call__(self, length: int, prior: torch.Tensor, temperature: float = 1.0, token_temperature: float = 1.0,
                 truncated_norm: Optional[Tuple[float, float]] = None, return_attn: bool = False,
                 encoder_temperature: float = 1.0, decode: bool = True, greedy: bool = False) -> Mapping:
        model = self.model
        u_c = model.encode_utterance(TensorMask(prior)) if self.has_utterance else None
        prior = model.encode(TensorMask(prior), temperature=encoder_temperature).value
        if not self.model_use_tokens:
            raise NotImplementedError("token-less LVTR variants are outside the VAE-GSLM hot path")
        model.transformer[0].cache_len_hint = prior.shape[1] + 1 + length + 2
        kw = dict(temperature=temperature, token_temperature=token_temperature, truncated_norm=truncated_norm,
                  greedy=greedy)
        it = model.step(prior, past_kv=None, return_attn=return_attn, push_init_state=True, **kw)
        state, kv = it["output"][:, -1:], it["kv"]
        frames = [prior, state]
        graphed = None
        for i in range(1, length):
            if self.use_cuda_graph and not return_attn and i >= 3 and prior.is_cuda:
                if graphed is None:
                    graphed = GraphedStep(model, state, kv, **kw)
                state = graphed().clone()
            else:
                it = model.step(state, past_kv=kv, return_attn=return_attn, **kw)
                state, kv = it["output"], it["kv"]
            frames.append(state)
        seq = torch.cat(frames, 1)
        outputs = {"frames": seq}
        if decode:
            outputs["output"] = model.decode(TensorMask(seq), u_c=u_c)
        return outputs
