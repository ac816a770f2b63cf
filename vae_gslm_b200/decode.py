"""Single-frame decode engine for the cached ``LVTR.step`` loop (reference ``models/speech/lvtr.py:227-286``,
``modules/transformer/layers.py:134-195`` with ``past_kv``).

The reference runs, per generated frame, ~40 small PyTorch kernels per layer and re-concatenates the KV cache.  One
frame per sequence is an HBM-bound problem (408.7 MB of bf16 weights + the KV cache per step), so this
engine runs the step as 5 kernels per layer, each using every SM:

    qkv   = decode_linear(x, RMSNorm1 folded in)                    attention.py:52  + norm.py:28-32
    o     = attn_decode(qkv, cache)   (appends k/v in place)        attention.py:56-85
    x     = decode_linear(o, out_proj, residual = x)  (+ Σx² out)   attention.py:79, transformer/layers.py:57
    h     = decode_linear(x, RMSNorm3 folded in, b1, GELU)          transformer/layers.py:82
    x     = decode_linear(h, linear2, b2, residual = x) (+ Σx² out) transformer/layers.py:82-86

``decode_linear`` streams a disjoint slab of the weight matrix per SM and prefetches it before the programmatic-
dependent-launch wait, so the weight stream of kernel i+1 overlaps the dependent tail of kernel i.  All buffers are
allocated once (CUDA-graph friendly); the row sum-of-squares that each RMSNorm needs is produced by the epilogue of the
GEMM that wrote the row.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU


# From this many sequences on the four linear layers of a transformer layer run on vg_skinny_linear (swap-AB tcgen05,
# cluster split-K through distributed shared memory) instead of vg_decode_linear (mma.sync): measured cross-over,
# profiles/r02_decode.md.
SKINNY_FROM_BATCH = 40


class DecodeEngine:
    def __init__(self, model, batch: int, device, overlap: bool = True, skinny: Optional[bool] = None) -> None:
        self.model, self.batch, self.device, self.overlap = model, batch, device, overlap
        self.skinny = (batch >= SKINNY_FROM_BATCH) if skinny is None else bool(skinny)
        stack = model.transformer[0]
        bf = torch.bfloat16
        assert stack.compute_dtype == bf, "the decode engine is the bf16 generation path"
        self.stack = stack
        self.nheads = stack.hp.layer.self_attn.nheads
        d = stack.hp.layer.dim
        self.dim = d
        f32 = lambda t: t.detach().float().contiguous()          # noqa: E731
        self.layers = []
        for lyr in stack.layers:
            if lyr._act_id is None or not lyr.preln:
                raise NotImplementedError("decode engine: pre-LN layers with a fused activation only")
            self.layers.append(dict(
                n1=f32(lyr.norm1.scale), eps1=lyr.norm1.eps, n3=f32(lyr.norm3.scale), eps3=lyr.norm3.eps,
                w_in=ops.lowp(lyr.self_attn.in_proj.weight, bf), w_out=ops.lowp(lyr.self_attn.out_proj.weight, bf),
                w1=ops.lowp(lyr.linear1.weight, bf), b1=f32(lyr.linear1.bias) if lyr.linear1.bias is not None else None,
                w2=ops.lowp(lyr.linear2.weight, bf), b2=f32(lyr.linear2.bias) if lyr.linear2.bias is not None else None,
                act=lyr._act_id))
            if self.skinny:
                # RMSNorm folded into the projection behind it: the norm's scale vector rides in the weight, 1/rms is applied
                # per row in the epilogue (norm.py:28-32 · attention.py:52 / transformer/layers.py:82)
                cur = self.layers[-1]
                cur["w_in_n"] = (lyr.self_attn.in_proj.weight.detach().float() * cur["n1"][None, :]).to(bf).contiguous()
                cur["w1_n"] = (lyr.linear1.weight.detach().float() * cur["n3"][None, :]).to(bf).contiguous()
        assert stack.linear is not None and stack.linear.bias is None and stack.final_norm is not None
        self.w_stack_in = ops.lowp(stack.linear.weight, bf)
        self.fn_scale, self.fn_eps = f32(stack.final_norm.scale), stack.final_norm.eps
        w_split, b_split = model._split_weights()
        w_head, b_head = model._head_weights()
        self.w_split, self.b_split = ops.lowp(w_split, bf), f32(b_split)
        self.w_head, self.b_head = ops.lowp(w_head, bf), f32(b_head)
        tp = model.token_predictor.linear
        self.w_tok, self.b_tok = ops.lowp(tp.weight, bf), f32(tp.bias)
        B = batch
        z = lambda *s, dt=bf: torch.zeros(*s, dtype=dt, device=device)       # noqa: E731
        ffd = self.layers[0]["w1"].shape[0]
        self.x, self.qkv, self.o, self.h = z(B, d), z(B, 3 * d), z(B, d), z(B, ffd)
        self.ss_a, self.ss_b = z(B, dt=torch.float32), z(B, dt=torch.float32)
        self.H = z(B, d)
        self.cg = z(B, self.w_split.shape[0])
        self.head = z(B, self.w_head.shape[0], dt=torch.float32)
        self.logits = z(B, self.w_tok.shape[0])
        self.ws = ops.decode_linear_workspace(B, max(3 * d, ffd, self.w_split.shape[0], self.w_head.shape[0],
                                                       self.w_tok.shape[0]), device)
        self.slopes = stack.rpe.slopes if stack.rpe is not None else None
        self.tickets = torch.zeros(B * self.nheads, dtype=torch.int32, device=device)    # in-kernel split-KV merge

    @torch.no_grad()
    def run(self, u: torch.Tensor, kv: List) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """u [B,64] fused input of the new frame; kv: the stack's LayerKV handles.  Returns (H, head, logits)."""
        B, ov, ws = self.batch, self.overlap, self.ws
        assert u.shape[0] == B
        cache = kv[0].cache
        cache.ensure(cache.length + 1)
        pos = cache.length
        u16 = u.to(torch.bfloat16).contiguous()
        # stack input linear (no bias); its epilogue produces Σx² for layer 0's RMSNorm.  Not overlapped: it must not
        # start before the previous step's last kernel is done with ss_a.
        self.ss_a.zero_()
        ops.decode_linear(u16, self.w_stack_in, ws, out=self.x, y_ss=self.ss_a, overlap=False)
        for i, lw in enumerate(self.layers):
            if self.skinny:
                ops.skinny_linear(self.x, lw["w_in_n"], x_ss=self.ss_a, norm_eps=lw["eps1"], out=self.qkv, zero_ss=self.ss_b)
                o = ops.attention_decode(self.qkv, cache.k(kv[i].index), cache.v(kv[i].index), pos, self.slopes,
                                         cache.pos_dev, out=self.o, tickets=self.tickets)
                ops.skinny_linear(o, lw["w_out"], residual=self.x, out=self.x, y_ss=self.ss_b)
                ops.skinny_linear(self.x, lw["w1_n"], lw["b1"], lw["act"], x_ss=self.ss_b, norm_eps=lw["eps3"], out=self.h,
                                  zero_ss=self.ss_a)
                ops.skinny_linear(self.h, lw["w2"], lw["b2"], residual=self.x, out=self.x, y_ss=self.ss_a)
                continue
            ops.decode_linear(self.x, lw["w_in"], ws, norm_scale=lw["n1"], x_ss=self.ss_a, norm_eps=lw["eps1"],
                              out=self.qkv, zero_ss=self.ss_b, overlap=ov)
            o = ops.attention_decode(self.qkv, cache.k(kv[i].index), cache.v(kv[i].index), pos, self.slopes,
                                     cache.pos_dev, out=self.o, tickets=self.tickets)
            ops.decode_linear(o, lw["w_out"], ws, residual=self.x, out=self.x, y_ss=self.ss_b, overlap=ov)
            ops.decode_linear(self.x, lw["w1"], ws, norm_scale=lw["n3"], x_ss=self.ss_b, norm_eps=lw["eps3"],
                              bias=lw["b1"], act=lw["act"], out=self.h, zero_ss=self.ss_a, overlap=ov)
            ops.decode_linear(self.h, lw["w2"], ws, bias=lw["b2"], residual=self.x, out=self.x, y_ss=self.ss_a,
                              overlap=ov)
        cache.length += 1
        if cache.pos_dev is not None:
            ops.L.call("vg_add_i32", ops.L.ptr(cache.pos_dev), 1, ops.L.stream())
        # final RMSNorm materialised (it is an output: 'transformer_latent'), then the three head GEMMs
        H = ops.rmsnorm(self.x, self.stack.final_norm.scale, self.fn_eps)
        self.H.copy_(H)
        d = self.dim
        ops.decode_linear(self.H, self.w_split, ws, bias=self.b_split, act=ACT_RELU, out=self.cg, overlap=False)
        ops.decode_linear(self.cg[:, :d], self.w_head, ws, bias=self.b_head, out_f32=self.head, overlap=ov)
        ops.decode_linear(self.cg[:, d:], self.w_tok, ws, bias=self.b_tok, out=self.logits, overlap=ov)
        return self.H, self.head, self.logits
