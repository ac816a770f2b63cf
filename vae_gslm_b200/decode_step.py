"""Persistent single-launch decode step (reference ``models/speech/lvtr.py:253-279``,
``modules/transformer/layers.py:134-195`` with ``past_kv``, ``modules/attention/attention.py:52-85``).

``vg_decode_step`` (csrc/decode_step.cu) interprets a task table: this module is the host side that builds it.

* the step is a list of phases — ``in`` (stack-input linear), per layer ``qkv | attn | out | ffn1 | ffn2``, then
  ``split`` (q_spliter | token_spliter) and ``heads`` (prior/FiLM head + logit head) — separated by device-wide barriers;
* every linear layer is cut into units of ``R`` output features x ``K / S`` inputs, one unit per CTA per phase; the
  decomposition is wide and short — up to 128 features x a 1/4 .. 1/16 k-slice — because a unit pays a full 128-row
  tcgen05 tile per k step whatever R is, and because the X operand a CTA must form shrinks with the k-slice;
* each CTA gets ONE contiguous byte stream holding the weight slabs of all its units in the order it will consume them,
  already in the shared-memory layout of the tcgen05 K-major SWIZZLE_128B operand, so the kernel's producer warp only
  bumps a pointer.

Interface-compatible with ``decode.DecodeEngine`` (``run(u, kv) -> (H, head, logits)``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU

GRID_MAX = 148
UNITS = 128                     # GEMM units per phase (power of two: every (R, S) below tiles N exactly or pads the tail)

TASK_DTYPE = np.dtype([
    ("x", "<u8"), ("acc", "<u8"), ("vec", "<u8"), ("ss_in", "<u8"), ("ss_out", "<u8"), ("bias_out", "<u8"),
    ("aux_ptr0", "<u8"), ("aux_ptr1", "<u8"), ("aux_ptr2", "<u8"), ("aux_ptr3", "<u8"),
    ("ldx", "<i4"), ("ldacc", "<i4"), ("k0", "<i4"), ("nkb", "<i4"), ("n0", "<i4"), ("rows", "<i4"), ("R", "<i4"),
    ("x_kind", "<i4"), ("act", "<i4"), ("store", "<i4"), ("inv_k", "<f4"), ("eps", "<f4"),
    ("aux_kind", "<i4"), ("aux_i0", "<i4"), ("aux_i1", "<i4"), ("aux_i2", "<i4"),
])


def _rows_per_unit(N: int, S: int, units: int = UNITS) -> int:
    """output features per unit (multiple of 16: the N of a tcgen05 M=128 tile) so that ceil(N / R) * S is about ``units``"""
    return max(16, (-(-(N * S) // units) + 15) // 16 * 16)


def split_factor(N: int, K: int, units: int = UNITS, max_split: int = 32, rmax: int = 128) -> Tuple[int, int]:
    """(R, S) for a [N, K] linear: the LARGEST k-split S (power of two) whose units still have R <= rmax output features.
    A tcgen05.mma of this kernel costs ~180 cycles whatever its N (measured, profiles/r02_decode.md), so the MMA time of a
    phase is proportional to the k-blocks per unit: wide-and-short units are the fast shape at EVERY batch size, and they
    keep the X operand a CTA must form small; the price is S partial sums per output, reduced with
    red.global.add.v4.f32 into the fp32 accumulators — which is why rmax is 256 only for small batches."""
    best = None
    S = 1
    while S <= max_split:
        R = _rows_per_unit(N, S, units)
        if R > rmax or (K // 64) % S or K // 64 // S < 1:
            break
        best = (R, S)
        S *= 2
    assert best is not None, (N, K)
    return best


def pack_units(W: torch.Tensor, R: int, S: int) -> torch.Tensor:
    """W [N, K] bf16 → [n_slabs, S, nkb * R * 64] bf16: per unit (feature slab, k-slice), per 64-wide k-block, R/8 atoms of
    8 rows x 128 bytes with the 16-byte groups of row r XOR-swizzled by r (SWIZZLE_128B, K-major) — exactly the bytes the
    tcgen05 descriptor expects at a 1024-byte aligned shared-memory address."""
    N, K = W.shape
    n_slabs = -(-N // R)
    nkb = K // 64 // S
    Wp = W
    if n_slabs * R != N:
        Wp = torch.zeros(n_slabs * R, K, dtype=W.dtype, device=W.device)
        Wp[:N] = W
    v = Wp.view(n_slabs, R // 8, 8, S, nkb, 8, 8)                  # slab, rg, r, s, kb, g, e
    v = v.permute(0, 3, 4, 1, 2, 5, 6)                             # slab, s, kb, rg, r, g, e
    r = torch.arange(8, device=W.device)
    src_g = (r[:, None] ^ r[None, :])                              # position g' of row r holds source group g' ^ r
    v = v[:, :, :, :, r[:, None], src_g, :]                        # slab, s, kb, rg, r, g', e
    return v.contiguous().view(n_slabs, S, nkb * R * 64)


def attention_splits(n_bh: int, n_warps: int, keys: int = 400) -> int:
    """kv-splits per (sequence, head) for the attention phases in warp-per-item mode: items = n_bh * splits are dealt
    round-robin to the ``n_warps`` worker warps of the grid.  Cost model fitted to the sweep of profiles/r02_decode.md
    (16 / 24 / 32 sequences x 1 / 2 / 4 / 8 splits, in units of one key per warp): every wave of items costs its keys plus
    ~40 of start-up (query, row statistics, first HBM round trip), and the merge by the last arriver grows with the split
    count:  t(s) = ceil(n_bh s / n_warps) (keys / s + 40) + 8 s."""
    best, best_t = 1, None
    for s in (1, 2, 3, 4, 6, 8, 12, 16):
        t = -(-n_bh * s // n_warps) * (keys / s + 40.0) + 8.0 * s
        if best_t is None or t < best_t:
            best, best_t = s, t
    return best


class _Args(C.Structure):
    _fields_ = [
        ("tasks", C.c_void_p), ("phase_kind", C.c_void_p), ("wstream", C.c_void_p), ("wstream_off", C.c_void_p),
        ("qkv_acc", C.c_void_p), ("ss_base", C.c_void_p), ("cache", C.c_void_p),
        ("cache_layer_stride", C.c_int64), ("cache_kv_stride", C.c_int64),
        ("attn_out", C.c_void_p), ("slopes", C.c_void_p), ("attn_partial", C.c_void_p), ("tickets", C.c_void_p),
        ("pos_dev", C.c_void_p), ("bar_flags", C.c_void_p), ("epoch", C.c_void_p), ("debug", C.c_void_p), ("trace", C.c_void_p),
        ("inv_d", C.c_float), ("eps", C.c_float), ("scale", C.c_float),
        ("NP", C.c_int32), ("grid", C.c_int32), ("B", C.c_int32), ("Bp", C.c_int32), ("H", C.c_int32),
        ("Tmax", C.c_int32), ("nsplit", C.c_int32), ("barrier_mode", C.c_int32), ("advance_pos", C.c_int32),
        ("rep", C.c_int32), ("attn_coop", C.c_int32), ("late_merge", C.c_int32),
    ]


class DecodeStepEngine:
    def __init__(self, model, batch: int, device, barrier_mode: int = -1, grid: int = 0, debug_phases: int = 0,
                 late_merge: Optional[bool] = None) -> None:
        lib = L.load()
        assert lib.vg_decode_step_task_bytes() == TASK_DTYPE.itemsize, "task struct layout drifted from include/vgslm.h"
        self.model, self.batch, self.device = model, batch, device
        stack = model.transformer[0]
        bf = torch.bfloat16
        assert stack.compute_dtype == bf, "the decode engine is the bf16 generation path"
        assert 1 <= batch <= 256
        self.stack = stack
        H = stack.hp.layer.self_attn.nheads
        d = stack.hp.layer.dim
        assert d == H * 64, "head dim 64"
        self.nheads, self.dim = H, d
        B = batch
        # rows of the X tile (the M operand): 128 or 256.  Up to 64 sequences the batch rows are REPLICATED into the other
        # TMEM lane quadrants (rows b + i * 128 / rep) so that all four epilogue warps hold the results and split the columns
        Bp = 128 if B <= 128 else 256
        self.Bp = Bp
        self.rep = 4 if B <= 32 else (2 if B <= 64 else 1)
        sms = torch.cuda.get_device_properties(device).multi_processor_count if torch.device(device).type == "cuda" else GRID_MAX
        G = grid or min(GRID_MAX, sms)
        assert G >= UNITS, "the unit decomposition assumes at least 128 SMs"
        self.grid = G
        f32 = lambda t: t.detach().float().contiguous()          # noqa: E731
        Lr = len(stack.layers)
        self.n_layers = Lr
        eps = float(stack.layers[0].norm1.eps)
        for lyr in stack.layers:
            if lyr._act_id != ACT_GELU or not lyr.preln:
                raise NotImplementedError("decode step kernel: pre-LN layers with GELU only")
            assert float(lyr.norm1.eps) == eps and float(lyr.norm3.eps) == eps
        assert stack.linear is not None and stack.linear.bias is None and stack.final_norm is not None
        assert float(stack.final_norm.eps) == eps
        ffd = stack.layers[0].linear1.weight.shape[0]
        din = stack.linear.weight.shape[1]
        assert din % 64 == 0
        w_split, b_split = model._split_weights()
        w_head, b_head = model._head_weights()
        tp = model.token_predictor.linear
        n_head, n_tok = w_head.shape[0], tp.weight.shape[0]
        self.n_head, self.n_tok = n_head, n_tok

        # ---- activations / accumulators (one zero-initialised fp32 region + the residual stream)
        def r4(n):
            return (n + 3) // 4 * 4
        off, o = {}, 0
        for name, n in (("ss", (2 * Lr + 1) * B), ("qkv", B * 3 * d), ("f1", B * ffd), ("cg", B * 2 * d),
                        ("head", B * n_head), ("logits", B * n_tok)):
            off[name] = o
            o += r4(n)
        self.scratch = torch.zeros(o, dtype=torch.float32, device=device)
        self.off = off
        sub = lambda name, n: self.scratch[off[name]: off[name] + n]          # noqa: E731
        self.ss = sub("ss", (2 * Lr + 1) * B)
        self.qkv_acc = sub("qkv", B * 3 * d)
        self.f1_acc = sub("f1", B * ffd)
        self.cg_acc = sub("cg", B * 2 * d)
        self.head = sub("head", B * n_head).view(B, n_head)
        self.logits = sub("logits", B * n_tok).view(B, n_tok)
        self.h = torch.zeros(B, d, dtype=torch.float32, device=device)
        self.o = torch.zeros(B, d, dtype=bf, device=device)
        self.H = torch.zeros(B, d, dtype=bf, device=device)
        self.u16 = torch.zeros(B, din, dtype=bf, device=device)
        # keep-alive list for the fp32 vectors the task table points at
        keep: List[torch.Tensor] = []

        def vec(t):
            t = f32(t)
            keep.append(t)
            return t.data_ptr()

        # ---- attention phases: how the (sequence, head, kv-split) items are dealt, and where their partials go
        # attention: few (sequence, head) pairs → one CTA per (pair, split), up to 4 kv-splits; many → one warp per (pair, split)
        self.attn_coop = 1 if B * H <= G else 0
        if os.environ.get("VG_DS_COOP"):                      # experiment switch
            self.attn_coop = int(os.environ["VG_DS_COOP"])
        self.nsplit = max(1, min(4, G // (B * H))) if self.attn_coop else attention_splits(B * H, G * 7)
        if os.environ.get("VG_DS_NSPLIT"):                    # experiment switch (tools/decode_bench.py sweeps)
            self.nsplit = int(os.environ["VG_DS_NSPLIT"])
        self.attn_partial = torch.zeros(B * H * self.nsplit * 72, dtype=torch.float32, device=device)
        self.tickets = torch.zeros(B * H, dtype=torch.int32, device=device)
        # late merge (default): the attention phases only publish (m, l, o) per kv-split; the out-projection's X transform
        # merges and normalises them (no ticket / last-arriver round trips on the critical path of the attention phase)
        self.late_merge = int(os.environ.get("VG_DS_LATE_MERGE", "1") != "0") if late_merge is None else int(late_merge)
        rmax = 256 if B <= 32 else 128          # output features per unit (= the N of the MMA): see split_factor
        # ---- phases
        rs_log: List[Tuple[int, int]] = []
        phases = []          # each: dict(kind=-1|layer, units=[...], aux=None|...)
        streams: List[List[torch.Tensor]] = [[] for _ in range(G)]
        inv_d = 1.0 / d

        def gemm_units(W, B_, x_ptr, ldx, x_kind, acc_ptr, ldacc, *, vec_ptr=0, act=ACT_NONE, ss_in=0, ss_out=0,
                       bias_out=0, store=0, inv_k=inv_d, n_off=0, max_split=32, units=UNITS):
            """units of one linear layer W [N, K] → list of (task dict, packed slab tensor)"""
            N, K = W.shape
            R, S = split_factor(N, K, units=units, max_split=1 if store else max_split, rmax=rmax)
            packed = pack_units(W.detach().to(bf), R, S)
            rs_log.append((R, S))
            n_slabs, nkb = packed.shape[0], K // 64 // S
            assert nkb & (nkb - 1) == 0, "k-blocks per unit must be a power of two"
            out = []
            for slab in range(n_slabs):
                for s in range(S):
                    t = dict(x=x_ptr, acc=acc_ptr, vec=vec_ptr, ss_in=ss_in,
                             ss_out=ss_out if slab == 0 else 0, bias_out=bias_out if s == 0 else 0,
                             ldx=ldx, ldacc=ldacc, k0=s * nkb * 64, nkb=nkb, n0=n_off + slab * R,
                             rows=min(R, N - slab * R), R=R, x_kind=x_kind, act=act, store=store, inv_k=inv_k, eps=eps)
                    out.append((t, packed[slab, s]))
            return out

        ss_ptr = lambda i: self.ss.data_ptr() + 4 * i * B          # noqa: E731
        zero_job = lambda name, n: ("zero", self.scratch.data_ptr(), off[name] // 4, r4(n) // 4)   # noqa: E731
        all_zero = ("zero", self.scratch.data_ptr(), 0, o // 4)

        # in: h = u · W_inᵀ (stored: K = one k-block, nothing to reduce); everything else is cleared alongside
        phases.append(dict(kind=-1, name="in", aux=all_zero, units=gemm_units(
            stack.linear.weight, B, self.u16.data_ptr(), din, 0, self.h.data_ptr(), d, store=1)))
        for i, lyr in enumerate(stack.layers):
            sa = lyr.self_attn
            phases.append(dict(kind=-1, name=f"qkv{i}", aux=zero_job("f1", B * ffd), units=gemm_units(
                sa.in_proj.weight, B, self.h.data_ptr(), d, 1, self.qkv_acc.data_ptr(), 3 * d, vec_ptr=vec(lyr.norm1.scale),
                ss_out=ss_ptr(2 * i))))
            phases.append(dict(kind=i, name=f"attn{i}", aux=None, units=[]))
            phases.append(dict(kind=-1, name=f"out{i}", aux=zero_job("qkv", B * 3 * d), units=gemm_units(
                sa.out_proj.weight, B, self.attn_partial.data_ptr() if self.late_merge else self.o.data_ptr(), d,
                3 if self.late_merge else 0, self.h.data_ptr(), d)))
            phases.append(dict(kind=-1, name=f"ffn1_{i}", aux=None, units=gemm_units(
                lyr.linear1.weight, B, self.h.data_ptr(), d, 1, self.f1_acc.data_ptr(), ffd, vec_ptr=vec(lyr.norm3.scale),
                ss_out=ss_ptr(2 * i + 1))))
            b1 = lyr.linear1.bias if lyr.linear1.bias is not None else torch.zeros(ffd, device=device)
            phases.append(dict(kind=-1, name=f"ffn2_{i}", aux=None, units=gemm_units(
                lyr.linear2.weight, B, self.f1_acc.data_ptr(), ffd, 2, self.h.data_ptr(), d, vec_ptr=vec(b1), act=ACT_GELU,
                ss_in=ss_ptr(2 * i + 1), bias_out=vec(lyr.linear2.bias) if lyr.linear2.bias is not None else 0)))
        fn = vec(stack.final_norm.scale)
        phases.append(dict(kind=-1, name="split", aux=None, units=gemm_units(
            w_split, B, self.h.data_ptr(), d, 1, self.cg_acc.data_ptr(), 2 * d, vec_ptr=fn, ss_out=ss_ptr(2 * Lr))))
        bs = f32(b_split)
        keep.append(bs)
        head_units = gemm_units(w_head, B, self.cg_acc.data_ptr(), 2 * d, 2, self.head.data_ptr(), n_head,
                                vec_ptr=bs.data_ptr(), act=ACT_RELU, ss_in=ss_ptr(2 * Lr), bias_out=vec(b_head),
                                units=96, max_split=8)
        tok_units = gemm_units(tp.weight, B, self.cg_acc.data_ptr() + 4 * d, 2 * d, 2, self.logits.data_ptr(), n_tok,
                               vec_ptr=bs.data_ptr() + 4 * d, act=ACT_RELU, ss_in=ss_ptr(2 * Lr), bias_out=vec(tp.bias),
                               units=40, max_split=8)
        phases.append(dict(kind=-1, name="heads", units=head_units + tok_units,
                           aux=("latent", self.h.data_ptr(), fn, ss_ptr(2 * Lr), self.H.data_ptr(), B * d // 8, d)))
        if debug_phases:                               # debugging aid: run only a prefix of the step
            phases = phases[:debug_phases]
        self.last_RS = sorted({(t["R"], t["nkb"]) for t, _ in phases[-1]["units"]})
        NP = len(phases)
        self.NP = NP
        self.phase_names = [p["name"] for p in phases]
        tasks = np.zeros((G, NP), dtype=TASK_DTYPE)
        tasks["eps"] = eps
        tasks["inv_k"] = inv_d
        n_bytes = 0
        for p, ph in enumerate(phases):
            units = ph["units"]
            assert len(units) <= G, (ph["name"], len(units))
            rot = (p * 20) % G                         # rotate the unit → CTA map so that all SMs stream weights
            for u, (t, slab) in enumerate(units):
                cta = (u + rot) % G
                for k, v in t.items():
                    tasks[cta, p][k] = v
                streams[cta].append(slab.reshape(-1))
                n_bytes += slab.numel() * 2
            aux = ph.get("aux")
            if aux is not None and aux[0] == "zero":
                _, base, i0, n4 = aux
                per = -(-n4 // G)
                for cta in range(G):
                    lo = min(n4, cta * per)
                    hi = min(n4, lo + per)
                    if hi > lo:
                        tasks[cta, p]["aux_kind"], tasks[cta, p]["aux_ptr0"] = 1, base
                        tasks[cta, p]["aux_i0"], tasks[cta, p]["aux_i1"] = i0 + lo, hi - lo
            elif aux is not None:
                _, hp_, sc_, ss_, out_, ngroups, dd = aux
                per = -(-ngroups // G)
                for cta in range(G):
                    lo = min(ngroups, cta * per)
                    hi = min(ngroups, lo + per)
                    if hi > lo:
                        tk = tasks[cta, p]
                        tk["aux_kind"], tk["aux_ptr0"], tk["aux_ptr1"], tk["aux_ptr2"], tk["aux_ptr3"] = 2, hp_, sc_, ss_, out_
                        tk["aux_i0"], tk["aux_i1"], tk["aux_i2"] = lo, hi - lo, dd
        self.weight_bytes = n_bytes
        offs, parts, o_b = [], [], 0
        for cta in range(G):
            offs.append(o_b)
            if streams[cta]:
                s_ = torch.cat(streams[cta])
                parts.append(s_)
                o_b += s_.numel() * 2
        self.wstream = torch.cat(parts) if parts else torch.zeros(8, dtype=bf, device=device)
        assert self.wstream.data_ptr() % 16 == 0 and all(x % 1024 == 0 for x in offs)
        self.wstream_off = torch.tensor(offs, dtype=torch.int64, device=device)
        self.tasks = torch.from_numpy(tasks.view(np.uint8).reshape(-1)).to(device)
        self.phase_kind = torch.tensor([p["kind"] for p in phases], dtype=torch.int32, device=device)
        self._keep = keep
        self.smem = lib.vg_decode_step_smem_bytes(NP)
        assert self.smem <= 227 * 1024, self.smem
        # ---- barrier state
        self.bar_flags = torch.zeros(G + 128, dtype=torch.int32, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.debug = torch.zeros(8, dtype=torch.int64, device=device)
        self.own_pos = torch.zeros(1, dtype=torch.int32, device=device)
        self.slopes = stack.rpe.slopes if stack.rpe is not None else None
        self.barrier_mode = barrier_mode if barrier_mode >= 0 else int(os.environ.get("VG_DS_BARRIER", "1"))
        self.eps = eps
        self.trace = None                     # set to a [512] int64 device tensor to record per-phase barrier clocks

    def _args(self, cache, layer0: int, pos_dev: torch.Tensor, advance: int) -> _Args:
        a = _Args()
        a.tasks, a.phase_kind = self.tasks.data_ptr(), self.phase_kind.data_ptr()
        a.wstream, a.wstream_off = self.wstream.data_ptr(), self.wstream_off.data_ptr()
        a.qkv_acc, a.ss_base = self.qkv_acc.data_ptr(), self.ss.data_ptr()
        buf = cache.buf
        assert buf.dtype == torch.bfloat16 and buf.is_contiguous() and buf.shape[-1] == 64
        a.cache = buf[layer0].data_ptr()
        a.cache_layer_stride, a.cache_kv_stride = buf.stride(0), buf.stride(1)
        a.attn_out = self.o.data_ptr()
        a.slopes = self.slopes.data_ptr() if self.slopes is not None else None
        a.attn_partial, a.tickets = self.attn_partial.data_ptr(), self.tickets.data_ptr()
        a.pos_dev, a.bar_flags, a.epoch, a.debug = (pos_dev.data_ptr(), self.bar_flags.data_ptr(), self.epoch.data_ptr(),
                                                    self.debug.data_ptr())
        a.trace = self.trace.data_ptr() if self.trace is not None else None
        a.inv_d, a.eps, a.scale = 1.0 / self.dim, self.eps, 1.0 / 8.0
        a.NP, a.grid, a.B, a.Bp, a.H = self.NP, self.grid, self.batch, self.Bp, self.nheads
        a.Tmax, a.nsplit, a.barrier_mode, a.advance_pos = cache.max_len, self.nsplit, self.barrier_mode, advance
        a.rep, a.attn_coop, a.late_merge = self.rep, self.attn_coop, self.late_merge
        return a

    @torch.no_grad()
    def run(self, u: torch.Tensor, kv: List) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """u [B, 64] fused input of the new frame; kv: the stack's LayerKV handles.  Returns (H bf16, head f32, logits f32)."""
        assert u.shape[0] == self.batch
        cache = kv[0].cache
        cache.ensure(cache.length + 1)
        self.u16.copy_(u)
        if cache.pos_dev is not None:              # CUDA-graph replay: the position lives on the device, the kernel advances it
            pos_dev, advance = cache.pos_dev, 1
        else:
            self.own_pos.fill_(cache.length)
            pos_dev, advance = self.own_pos, 0
        a = self._args(cache, kv[0].index, pos_dev, advance)
        L.call("vg_decode_step", C.byref(a), L.stream())
        cache.length += 1
        return self.H, self.head, self.logits
