"""VAE-GSLM model (reference ``models/speech/lvtr.py:18-395``), B200-native.

Same constructor, sub-module names / state-dict keys and public methods as the reference ``LVTR``;
``forward`` and ``step`` run on the fused CUDA kernels of libvgslm:

    conv encoder (dwconv_ln kernel + 1x1-conv GEMMs) → latent_front kernel → stack-input GEMM → 16 × [RMSNorm, QKV
    GEMM, attention, out-proj GEMM(+res+mask), RMSNorm, FFN1 GEMM(+bias+GELU), FFN2 GEMM(+bias+res+mask)]
    → RMSNorm → one GEMM for q_spliter|token_spliter (+bias+ReLU) → one GEMM for prior mean|logstd|4×FiLM
    → latent_back kernel (flow + log_p + KL) ; logit GEMM → softmax-CE kernel ; diffusion UNet (same conv kernels)
    between the q_sample and masked-L1 kernels, on a side stream concurrently with the transformer.
    Single-frame cached steps (``step``) run on ``decode.DecodeEngine`` (weight-streaming ``vg_decode_linear``).

Additive API: the five RNG draws of the reference forward can be injected
(``eps_q, init_state, eps_p, diff_t, diff_noise``), ``step`` accepts ``eps=`` / ``token_u=`` /
``greedy=True``; ``set_compute_dtype(torch.bfloat16)`` switches the activation stream to bf16.
"""
from __future__ import annotations

import os

import contextlib
import math
from typing import List, Mapping, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...hparams.hp import Hparams
from ...modules.conv.layers import BottleNeckResNet, CNNStack
from ...modules.diffusion.ddpm import GaussianDiffusion1D
from ...modules.diffusion.unet import ConditionalBottleNeckUNet
from ...modules.flow.layers import CouplingStack
from ...modules.linear.layers import Embedding, GaussianParameterize, Linear, TimeAggregation
from ...modules.transformer.layers import TransformerLayerStack
from ...utils.attr import AttrDict
from ...utils.tensormask import TensorMask


class LVTR(nn.Module):
    def __init__(self, hp: Hparams, input_dim: Optional[int] = None, memory_dim: Optional[int] = None) -> None:
        super().__init__()
        hp.check_arg_in_hparams("encoder", "decoder", "transformer", "latent_dim")
        self.input_dim = input_dim
        self.hp = hp
        enc_kind = hp.encoder.get("identifier", "ResNet")
        if enc_kind == "BottleNeckResNet":
            encoder_model = BottleNeckResNet
        elif enc_kind == "CNNStack":
            encoder_model = CNNStack
        else:
            raise NotImplementedError(f"encoder {enc_kind}: only BottleNeckResNet / CNNStack are built")
        self.encoder = nn.Sequential(
            encoder_model(hp.encoder, input_dim=input_dim, output_dim=hp.latent_dim),
            GaussianParameterize(hp.latent_dim, hp.latent_dim,
                                 std=hp.encoder.get("fix_std", None), std_range=hp.encoder.get("std_range", None),
                                 truncated_norm=hp.encoder.get("truncated_norm", None),
                                 total_std=hp.encoder.get("total_std", None), use_tanh=False,
                                 normalization=hp.encoder.get("normalization", False)))
        self.tokens = hp.get("tokens", None)
        dim = hp.transformer.layer.dim
        if self.tokens is not None:
            self.tokens.check_arg_in_hparams("embedding_dim", "vocab_size")
            self.token_embedding_dim = self.tokens.embedding_dim
            self.token_embedding = Embedding(self.tokens.vocab_size, self.tokens.embedding_dim)
            self.token_predictor = Linear(dim, self.tokens.vocab_size)
            self.token_fuser = Linear(hp.latent_dim, self.tokens.embedding_dim, activation=nn.ReLU())
            self.token_spliter = Linear(dim, dim, activation=nn.ReLU())
            self.q_spliter = Linear(dim, dim, activation=nn.ReLU())
        else:
            self.q_spliter = nn.Identity()
        diff_cond_dim = self.tokens.embedding_dim if self.tokens is not None else hp.latent_dim
        if hp.has("utterance_encoder"):
            diff_cond_dim += hp.utterance_encoder.embedding_dim
        dec_kind = hp.decoder.diffusion.get("identifier", "ConditionalUNet")
        if dec_kind != "ConditionalBottleNeckUNet":
            raise NotImplementedError(f"decoder {dec_kind}: only ConditionalBottleNeckUNet is built")
        hp.decoder.check_arg_in_hparams("cond_unet")
        self.decoder = GaussianDiffusion1D(ConditionalBottleNeckUNet(diff_cond_dim, input_dim, hp.decoder.cond_unet),
                                           hp.decoder.diffusion)
        self.diff_scaling = hp.decoder.diffusion.get("input_scale", 1.0)
        self.transformer_flow = None
        if hp.transformer.has("flow"):
            cond_dim = dim if hp.transformer.flow.get("conditional", False) else None
            self.transformer_flow = CouplingStack(hp.latent_dim, hp.transformer.flow, condition_dim=cond_dim)
        tr_in = self.tokens.embedding_dim if self.tokens is not None else hp.latent_dim
        self.transformer = nn.Sequential(
            TransformerLayerStack(hp.transformer, input_dim=tr_in, memory_dim=memory_dim),
            GaussianParameterize(dim, hp.latent_dim, std=hp.transformer.get("fix_std", None),
                                 std_range=hp.transformer.get("std_range", None), use_tanh=False,
                                 mean=hp.transformer.get("fix_mean", None)))
        self.utterance_encoder = None
        if hp.has("utterance_encoder"):
            self.utterance_encoder = nn.Sequential(
                CNNStack(hp.utterance_encoder, input_dim=input_dim, output_dim=hp.utterance_encoder.embedding_dim),
                TimeAggregation())
        self.use_tokens = self.tokens is not None
        self.compute_dtype = torch.float32
        self.overlap_decoder = True         # diffusion decoder branch on a side stream (parallel graph branch)
        self.use_decode_engine = True       # bf16 single-frame steps run on decode.DecodeEngine ...
        self.decode_pdl = os.environ.get("VG_DECODE_PDL", "1") != "0"     # layer-by-layer cached step as a PDL chain
        # ... up to this batch: measured on B200 (profiles/r01_decode.md) the weight-streaming engine wins up to ~100
        # sequences (0.47 vs 1.21 ms per step at B=1, 1.02 vs 1.37 at 64), the tcgen05 layer-by-layer path above
        self.decode_engine_max_batch = 96
        # Which engine runs a bf16 single-frame step.  Measured on B200 over the configured generation (3 s prompt →
        # 10 s, mean 402 cached keys; profiles/r02_decode.md), ms per step for step / linear / layer-by-layer:
        #   batch 1: 0.47 / 0.54 / 1.2     8: 0.53 / 0.74 / -     32: 0.85 / 0.89 / -     64: 1.41 / 1.24 / 1.42
        #   128: 2.33 / - / 1.71     256: 4.27 / - / 2.26
        # "auto" = the persistent single-launch kernel (decode_step.py) up to 32 sequences, round 1's kernel-per-linear
        # engine (decode.py) up to 80, the tcgen05 layer-by-layer path above.  "step" / "linear" force one engine.
        self.decode_engine_kind = "auto"

    # ------------------------------------------------------------------ configuration
    def set_compute_dtype(self, dtype: torch.dtype) -> "LVTR":
        """float32 = parity mode (fp32 CUDA-core GEMMs); bfloat16 = tcgen05 GEMMs, bf16 activations."""
        assert dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = dtype
        self.transformer[0].compute_dtype = dtype
        nets = [self.encoder[0], self.decoder.model.unet]                 # conv stacks on libvgslm (B,T,C layout)
        if self.utterance_encoder is not None:
            nets.append(self.utterance_encoder[0])
        for net in nets:
            if hasattr(net, "compute_dtype"):
                net.compute_dtype = dtype
        return self

    def _autocast(self):
        if self.compute_dtype == torch.bfloat16:
            return torch.autocast("cuda", dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def _check_fused(self) -> None:
        ok = (self.use_tokens and self.encoder[1].is_plain and self.transformer[1].is_plain
              and self.transformer_flow is not None and self.transformer_flow.fusable
              and self.hp.latent_dim == 4 and self.token_embedding_dim == 64
              and isinstance(self.token_fuser.activation, nn.ReLU))
        if not ok:
            raise NotImplementedError(
                "this configuration is outside the fused VAE-GSLM hot path (tokens + plain Gaussian heads + "
                "4-layer conditional LinearCoupling flow, latent_dim=4, embedding_dim=64)")

    @property
    def sample_ratio(self) -> float:
        return self.encoder[0].sample_ratio

    # ------------------------------------------------------------------ shared pieces
    def _cat(self, name: str, tensors, dim: int = 0) -> torch.Tensor:
        """concatenate parameters for a fused GEMM.  With autograd on this is a differentiable torch.cat (the
        gradient splits back to the individual parameters); in inference the result is memoised until a
        source parameter changes, so the decode loop does not re-concatenate (or re-cast) every step."""
        if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
            return torch.cat(list(tensors), dim)
        key = (self._weights_generation(),) + tuple((t.data_ptr(), t._version) for t in tensors)
        memo = self.__dict__.setdefault("_cat_memo", {})
        hit = memo.get(name)
        if hit is None or hit[0] != key:
            hit = (key, torch.cat([t.detach() for t in tensors], dim))
            memo[name] = hit
        return hit[1]

    def _weights_generation(self) -> int:
        """counter of parameter updates torch cannot see (ParamArena's fused AdamW writes through raw pointers and leaves
        Tensor._version alone): part of every key under which derived weights are memoised"""
        arena = self.__dict__.get("_vg_param_arena")
        return arena.generation if arena is not None else 0

    def _split_weights(self):
        wq, wt = self.q_spliter.linear, self.token_spliter.linear
        return self._cat("split_w", (wq.weight, wt.weight)), self._cat("split_b", (wq.bias, wt.bias))

    def _head_weights(self):
        prior = self.transformer[1]
        films = [l.film.linear for l in self.transformer_flow.layers]
        return (self._cat("head_w", (prior.mean.weight, prior.logstd.weight, *[f.weight for f in films])),
                self._cat("head_b", (prior.mean.bias, prior.logstd.bias, *[f.bias for f in films])))

    def _post_stack(self, H: torch.Tensor):
        """transformer_latent → (flow condition c, prior/FiLM head [.,520] f32, logits)."""
        D = H.shape[-1]
        w_split, b_split = self._split_weights()
        cg = ops.linear(H, w_split, b_split, act=ops.ACT_RELU)                       # q_spliter | token_spliter
        c_lat, g_lat = cg[..., :D], cg[..., D:]
        w_head, b_head = self._head_weights()
        head = ops.linear(c_lat, w_head, b_head, out_dtype=torch.float32)
        logits = ops.linear(g_lat, self.token_predictor.linear.weight, self.token_predictor.linear.bias)
        return c_lat, head, logits

    # ------------------------------------------------------------------ training forward
    def forward(self, x: TensorMask, c: Optional[TensorMask] = None, spkr: Optional[torch.Tensor] = None,
                utterance: Optional[TensorMask] = None, diff_input: Optional[TensorMask] = None, *,
                eps_q: Optional[torch.Tensor] = None, init_state: Optional[torch.Tensor] = None,
                eps_p: Optional[torch.Tensor] = None, diff_t: Optional[torch.Tensor] = None,
                diff_noise: Optional[torch.Tensor] = None) -> Mapping[str, TensorMask]:
        self._check_fused()
        if diff_input is not None:
            raise NotImplementedError("diff_input (separate diffusion target) is not used by the VAE-GSLM recipe")
        mask = x.mask
        tokens_id, mel = self.split_inputs(x)
        tokens_id = tokens_id.long().squeeze(-1)
        B, T = mask.shape
        dev = mel.value.device
        Ld = self.hp.latent_dim
        # The utterance encoder (strided convs over <= 200 frames: ~100 tiny launches fwd + bwd) does not meet the
        # main path until the diffusion decoder: run it on a side stream (a parallel branch of the captured graph;
        # autograd replays its backward on the same stream) so that it hides behind the transformer GEMMs.
        u_c, side = None, None
        if self.utterance_encoder is not None:
            if mel.value.is_cuda:
                side = self.__dict__.setdefault("_side_stream", torch.cuda.Stream(device=dev))
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side), self._autocast():
                    u_c = self.utterance_encoder(utterance)
            else:
                u_c = self.utterance_encoder(utterance)
        with self._autocast():
            h_enc = self.encoder[0](mel).value.float()
        # RNG draws #1..#3 in the reference's order (lvtr.py:156,161,172)
        if eps_q is None:
            eps_q = torch.randn(B, T, Ld, device=dev)
        if init_state is None:
            init_state = self.initial_state(B, dev)
        if eps_p is None:
            eps_p = torch.randn(B, T, Ld, device=dev)      # drawn by the reference's prior head, unused by the loss
        post, fuser = self.encoder[1], self.token_fuser.linear
        mean_q, logstd_q, z, log_q, u, u_shift = ops.latent_front(
            h_enc, eps_q, tokens_id.value, mask, init_state.reshape(B, -1), post.mean.weight, post.mean.bias,
            post.logstd.weight, post.logstd.bias, self.token_embedding.weight, fuser.weight, fuser.bias,
            1.0, self.compute_dtype)
        # diffusion decoder on fuse(z, tokens) ⊕ utterance embedding (lvtr.py:201-209).  It depends on the posterior
        # sample only, not on the transformer: with `overlap_decoder` it runs on the side stream, concurrently with the
        # transformer stack (a parallel branch of the captured graph, forward and backward) — its many short kernels
        # fill the gaps of the big GEMMs instead of extending the critical path.
        def run_decoder():
            cond = u
            if u_c is not None:
                cond = torch.cat([cond, u_c.to(cond.dtype)[:, None].expand(-1, T, -1)], -1)
            with self._autocast():
                return self.decoder(mel / self.diff_scaling, TensorMask(cond, mask), t=diff_t, noise=diff_noise)

        rec_x = None
        main = torch.cuda.current_stream() if side is not None else None
        if side is not None and self.overlap_decoder:
            side.wait_stream(main)                       # u (and the RNG draws above) are ordered before the branch
            for t_ in (u, mel.value, mask):
                t_.record_stream(side)
            with torch.cuda.stream(side):
                rec_x = run_decoder()
        transformer_latent = self.transformer[0](TensorMask(u_shift, mask), c)
        c_lat, head, logits = self._post_stack(transformer_latent.value)
        flow = self.transformer_flow
        lo, hi = flow.scale_lo_hi
        log_p, sample_p, kl_sum = ops.latent_back(head, z, log_q, mask, *flow.stacked_parameters(), flow.ln_eps, lo, hi)
        ce_loss = ops.softmax_ce(logits, tokens_id.value, mask)
        if side is not None:
            main.wait_stream(side)
            if u_c is not None:
                u_c.record_stream(main)
            if rec_x is not None:
                (rec_x.value if isinstance(rec_x, TensorMask) else rec_x).record_stream(main)
        if rec_x is None:
            rec_x = run_decoder()
        q_z = AttrDict(mean=TensorMask(mean_q, mask), logstd=TensorMask(logstd_q, mask), sample=TensorMask(z, mask))
        with torch.no_grad():
            hd = head.detach()
            prior_mean, prior_logstd = TensorMask(hd[..., :Ld], mask), TensorMask(hd[..., Ld:2 * Ld], mask)
            stats = dict(logstd=prior_logstd.mean(), mean=prior_mean.mean(),
                         q_logstd=TensorMask(logstd_q.detach(), mask).mean(),
                         q_mean=TensorMask(mean_q.detach(), mask).mean(),
                         q_mean_abs=TensorMask(mean_q.detach().abs(), mask).mean())
        return {
            "log_p": TensorMask(log_p, mask),
            "log_q": TensorMask(log_q, mask),
            "decoder_output": rec_x,
            "sample_q": TensorMask(z, mask),
            "transformer_latent": transformer_latent,
            "q_z": q_z,
            "u_c": u_c,
            "ce_loss": ce_loss,
            "kl_sum": kl_sum,            # fused Σ_valid mean_c(log_q − log_p) (trainers/speech/lvtr.py:122-124)
            "logits": TensorMask(logits, mask),
            **stats,
        }

    # ------------------------------------------------------------------ cached generation step
    @torch.no_grad()
    def step(self, x: torch.Tensor, c: Optional[TensorMask] = None, spkr: Optional[torch.Tensor] = None,
             past_kv: Optional[List] = None, temperature: float = 1.0, token_temperature: float = 1.0,
             truncated_norm: Optional[Tuple[float, float]] = None, return_attn: bool = False,
             return_distrbution: bool = False, push_init_state: bool = False, *,
             eps: Optional[torch.Tensor] = None, token_u: Optional[torch.Tensor] = None, greedy: bool = False,
             init_state: Optional[torch.Tensor] = None, return_logits: bool = False, **kwargs) -> Mapping:
        """x: [B, t, 1 + latent] = (token id as float, z).  Returns the reference's dict (+ optional logits)."""
        self._check_fused()
        ids = x[..., 0].long()
        zin = x[..., 1:].float()
        fuser = self.token_fuser.linear
        u = F.embedding(ids, self.token_embedding.weight) + F.relu(F.linear(zin, fuser.weight, fuser.bias))
        if push_init_state:
            s0 = self.initial_state(x.shape[0], x.device) if init_state is None else init_state
            u = torch.cat([s0.reshape(x.shape[0], 1, -1).to(u.dtype), u], 1)
        stack = self.transformer[0]
        engine = self._decode_engine(u, past_kv, c, return_attn, return_distrbution)
        if engine is not None:
            # single new frame per sequence, bf16: the weight-streaming decode engine (decode.py), 5 kernels per layer
            H2, head2, logits2 = engine.run(u[:, 0], past_kv)
            H = H2.view(u.shape[0], 1, -1)
            head, logits = head2.view(u.shape[0], 1, -1), logits2.view(u.shape[0], 1, -1)
            outputs = {"transformer_latent": TensorMask(H), "kv": past_kv}
        else:
            # a single cached frame in bf16 is ~120 short launches (7 per layer): run them as a programmatic-dependent-
            # launch chain with the weights prefetched ahead of each dependency wait (ops.pdl_mode)
            chain = (u.shape[1] == 1 and past_kv is not None and u.is_cuda and self.compute_dtype == torch.bfloat16
                     and not torch.is_grad_enabled() and self.decode_pdl)
            with ops.pdl_mode(3 if chain else 0):
                z_given = stack.run(TensorMask(u), memory=c, past_kv=past_kv, return_attn=return_attn, return_kv=True)
                H = z_given["output"].value
                c_lat, head, logits = self._post_stack(H)
            outputs = {"transformer_latent": z_given["output"], "kv": z_given["kv"]}
            if return_distrbution:
                outputs["z_given"] = z_given
            if return_attn:
                outputs["self_attn"] = z_given["self_attn"]
        Bq, Tq, _ = H.shape
        Ld = self.hp.latent_dim
        if eps is None:
            eps = torch.randn(Bq, Tq, Ld, device=H.device)
            if truncated_norm is not None:
                nn.init.trunc_normal_(eps, a=truncated_norm[0], b=truncated_norm[1])
        flow = self.transformer_flow
        lo, hi = flow.scale_lo_hi
        z_new = ops.latent_prior_sample(head, eps, temperature, *flow.stacked_parameters(), flow.ln_eps, lo, hi)
        if greedy:
            tok = ops.sample_token(logits, None)
        else:
            if token_u is None:
                token_u = torch.rand(Bq * Tq, device=H.device)
            tok = ops.sample_token(logits, token_u, token_temperature)
        outputs["output"] = torch.cat([tok.reshape(Bq, Tq, 1).float(), z_new], -1)
        if return_logits:
            outputs["logits"] = logits
        return outputs

    def _decode_engine(self, u: torch.Tensor, past_kv, c, return_attn: bool, return_distribution: bool):
        """the DecodeEngine for this batch size when the call is a plain single-frame cached bf16 step, else None."""
        from ...modules.attention.kvcache import LayerKV
        if (not self.use_decode_engine or u.shape[1] != 1 or c is not None or return_attn or return_distribution
                or past_kv is None or not isinstance(past_kv[0], LayerKV) or not u.is_cuda
                or self.compute_dtype != torch.bfloat16):
            return None
        nb = u.shape[0]
        kind = self.decode_engine_kind
        if kind == "auto":
            kind = "step" if nb <= 20 else ("linear" if nb <= 64 else "none")      # measured cross-overs: profiles/r02_decode.md
        step_kind = kind == "step"
        if kind == "none" or nb > (256 if step_kind else self.decode_engine_max_batch):
            return None
        engines = self.__dict__.setdefault("_decode_engines", {})
        stack = self.transformer[0]
        # rebuilt after a weight update, seen by torch (_version) or not (the arena's generation counter)
        stamp = (kind, self._weights_generation()) + tuple(p._version for p in stack.parameters())
        ent = engines.get(u.shape[0])
        if ent is None or ent[0] != stamp:
            if step_kind:
                from ...decode_step import DecodeStepEngine
                ent = (stamp, DecodeStepEngine(self, u.shape[0], u.device))
            else:
                from ...decode import DecodeEngine
                ent = (stamp, DecodeEngine(self, u.shape[0], u.device, skinny=self.__dict__.get("decode_engine_skinny")))
            engines[u.shape[0]] = ent
        return ent[1]

    # ------------------------------------------------------------------ auxiliary entry points
    @torch.no_grad()
    def decode(self, x: TensorMask, c: Optional[TensorMask] = None, u_c: Optional[torch.Tensor] = None, *,
               start_noise: Optional[torch.Tensor] = None, step_noise=None) -> TensorMask:
        """diffusion sampling of mel frames from (token, z) frames (lvtr.py:288-306): the sampling loop
        is host code, every UNet pass runs on the libvgslm conv / GEMM kernels.  ``start_noise`` / ``step_noise``
        inject the reference's torch.randn / torch.randn_like draws (additive API, parity tests)."""
        n_frames = int(x.value.size(1) * (1.0 / self.sample_ratio))
        noise = (start_noise.to(x.value.device) if start_noise is not None
                 else torch.randn(x.value.size(0), n_frames, self.input_dim, device=x.device))
        noise = TensorMask.fromlength(noise, TensorMask.resize_length(x.length, 1.0 / self.sample_ratio)).apply_mask()
        if self.use_tokens:
            tokens_id, x = x.split(1)
            tokens = self.token_embedding(tokens_id.long().squeeze(-1))
            x = self.fuse_inputs(x, tokens)
        if u_c is not None:
            x = x.cat(u_c[:, None].expand(-1, x.value.size(1), -1).to(x.value.dtype))
        with self._autocast():
            return self.decoder.sample(noise, x.apply_mask(), step_noise=step_noise) * self.diff_scaling

    @torch.no_grad()
    def encode(self, x: TensorMask, temperature: float = 1.0, beta: Optional[torch.Tensor] = None,
               utterance: Optional[TensorMask] = None, eps: Optional[torch.Tensor] = None) -> TensorMask:
        if self.use_tokens:
            tokens_id, x = self.split_inputs(x)
        if beta is not None:
            raise NotImplementedError("beta conditioning is stale code in the reference (SURVEY §3.3)")
        with self._autocast():
            h = self.encoder[0](x)
        out = self.encoder[1](TensorMask(h.value.float(), h.mask), temperature, eps=eps).sample
        if self.use_tokens:
            return tokens_id.cat(out.apply_mask())
        return out.apply_mask()

    @torch.no_grad()
    def encode_utterance(self, utterance: TensorMask) -> torch.Tensor:
        if self.use_tokens:
            _, utterance = self.split_inputs(utterance)
        with self._autocast():
            return self.utterance_encoder(utterance)

    def initial_state(self, bsize: int, device=None, nfeat: Optional[int] = None) -> torch.Tensor:
        """random BOS state ~ U(−1, 1) [B,1,nfeat] (lvtr.py:328-335) — RNG draw #2."""
        if nfeat is None:
            nfeat = self.token_embedding_dim if self.tokens is not None else self.hp.latent_dim
        return torch.rand(bsize, 1, nfeat, device=device) * 2.0 - 1.0

    @torch.no_grad()
    def likelihood(self, x: TensorMask, temperature: float = 0.0, gamma: Optional[float] = 1.0,
                   init_state: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        """per-utterance mean token log-probability with a temperature-0 posterior (lvtr.py:337-388)."""
        self._check_fused()
        mask = x.mask
        tokens_id, mel = self.split_inputs(x)
        ids = tokens_id.long().squeeze(-1).value
        B, T = mask.shape
        with self._autocast():
            h_enc = self.encoder[0](mel).value.float()
        eps = torch.randn(B, T, self.hp.latent_dim, device=h_enc.device)
        s0 = self.initial_state(B, h_enc.device) if init_state is None else init_state
        post, fuser = self.encoder[1], self.token_fuser.linear
        _, _, _, _, _, u_shift = ops.latent_front(
            h_enc, eps, ids, mask, s0.reshape(B, -1), post.mean.weight, post.mean.bias, post.logstd.weight,
            post.logstd.bias, self.token_embedding.weight, fuser.weight, fuser.bias, float(temperature),
            self.compute_dtype)
        H = self.transformer[0](TensorMask(u_shift, mask))
        _, _, logits = self._post_stack(H.value)
        logp = torch.log_softmax(logits.float(), -1).gather(-1, ids.unsqueeze(-1)).squeeze(-1)
        return TensorMask.use_mask(logp, mask).sum(-1) / mask.sum(-1)

    def fuse_inputs(self, x: TensorMask, tokens: TensorMask) -> TensorMask:
        return tokens + self.token_fuser(x)

    def split_inputs(self, x: TensorMask):
        return x.split(1)
