"""Flat parameter / gradient / optimizer-state arenas (B200-first memory layout for the training step).

All parameters of the model are re-pointed at views of two contiguous fp32 buffers (weight-decayed
tensors | 1-D tensors, the split of training_lib/optimizer.py:110-125), in forward execution order.
Gradients, AdamW moments and the bf16 shadow weights mirror that layout, so that

  * the optimizer is TWO launches of the fused AdamW kernel (which also refreshes the bf16 shadows),
    instead of torch's multi-tensor AdamW + 100 cast kernels;
  * data-parallel gradient buckets are contiguous slices (no flatten/unflatten copies) that become
    ready in roughly reverse order during backward (dp.py);
  * the wgrad GEMMs of the big transformer matrices write straight into the gradient arena
    (``param._vg_main_grad``) instead of going through autograd's accumulate pass.

state_dict() / load_state_dict() keep working: parameters are ordinary nn.Parameters whose storage
happens to live in the arena.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L

_ALIGN = 64      # elements; keeps every tensor 256-byte aligned inside the arena (TMA needs 16 B)


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def execution_order(model: nn.Module) -> List[Tuple[str, nn.Parameter]]:
    """parameters in (approximate) forward execution order, so that backward completes buckets back to front."""
    prefixes = ["encoder.", "token_embedding.", "token_fuser.", "transformer.0.", "q_spliter.", "token_spliter.",
                "transformer.1.", "transformer_flow.", "token_predictor.", "utterance_encoder.", "decoder."]
    named = list(model.named_parameters())
    out, seen = [], set()
    for pre in prefixes:
        for n, p in named:
            if n.startswith(pre) and n not in seen:
                out.append((n, p))
                seen.add(n)
    out += [(n, p) for n, p in named if n not in seen]
    return out


class _Group:
    """one contiguous arena: params, grads, exp_avg, exp_avg_sq (+ optional bf16 shadow)."""

    def __init__(self, items: List[Tuple[str, nn.Parameter]], device, weight_decay: float, shadow: bool):
        self.names = [n for n, _ in items]
        self.params = [p for _, p in items]
        self.weight_decay = weight_decay
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = max(off, _ALIGN)
        f32 = dict(dtype=torch.float32, device=device)
        self.p = torch.zeros(self.numel, **f32)
        self.g = torch.zeros(self.numel, **f32)
        self.m = torch.zeros(self.numel, **f32)
        self.v = torch.zeros(self.numel, **f32)
        self.shadow = torch.zeros(self.numel, dtype=torch.bfloat16, device=device) if shadow else None
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.p[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.p[o:o + n].view(p.shape)
            p.grad = self.g[o:o + n].view(p.shape)
            if shadow and p.dim() >= 2:
                p._vg_shadow = self.shadow[o:o + n].view(p.shape)
        if shadow:
            L.call("vg_cast_f32_to_bf16", L.ptr(self.p), L.ptr(self.shadow), self.numel, L.stream())

    def slice_of(self, index: int) -> Tuple[int, int]:
        return self.offsets[index], self.params[index].numel()


class ParamArena:
    def __init__(self, model: nn.Module, weight_decay: float = 0.1, bf16_shadow: bool = True,
                 direct_wgrad: bool = True) -> None:
        device = next(model.parameters()).device
        items = execution_order(model)
        self.decay = _Group([(n, p) for n, p in items if p.ndim != 1], device, weight_decay, bf16_shadow)
        self.nodecay = _Group([(n, p) for n, p in items if p.ndim == 1], device, 0.0, False)
        self.groups = [self.decay, self.nodecay]
        self.step_count = 0
        # bumped whenever parameter VALUES change behind torch's back (the fused AdamW writes through raw pointers and
        # does not touch Tensor._version): inference-side memos of derived weights (LVTR._cat, the decode engines'
        # packed / concatenated copies) key on it
        self.generation = 0
        model.__dict__["_vg_param_arena"] = self
        self.micro_batch = 0          # index inside the current accumulation window (0 → wgrad overwrites)
        self.hyper = [torch.zeros(4, dtype=torch.float32, device=device) for _ in self.groups]
        # pinned staging of {lr, 1/bc1, 1/sqrt(bc2), 1-lr*wd}: a ring of _HYPER_DEPTH buffers per group, one per step, so
        # that a host running ahead of the GPU never overwrites values whose (stream-ordered, asynchronous) upload has not
        # executed yet; upload_hyper() guards the reuse of a ring entry with an event
        pin = torch.cuda.is_available()
        self._hyper_ring = [[torch.zeros(4, dtype=torch.float32).pin_memory() if pin else torch.zeros(4)
                             for _ in self.groups] for _ in range(self._HYPER_DEPTH)]
        self._hyper_events = [None] * self._HYPER_DEPTH
        self._hyper_host = self._hyper_ring[0]
        self.on_grad_ready = None     # dp.py installs a callback(param) here
        # Every parameter carries a handle to its gradient slice: the ops that produce parameter gradients on the hot
        # path (wgrad GEMMs, bias column sums, RMSNorm scale sums — ops._wgrad / ops._bgrad / ops._RMSNorm) ACCUMULATE
        # into it directly (beta = 1) and hand autograd None, so no AccumulateGrad add kernel runs for them.
        # Gradients that still come back through autograd (torch ops, the fused latent / conv kernels) are added into
        # the same slice by autograd itself.  Both are accumulations, hence one whole-arena clear per step.
        #
        # After one CALIBRATION step (calibrate_next → finish_calibration) the arena knows which parameters are written
        # ONLY through those direct sites.  From then on the first direct write of a step OVERWRITES (beta = 0), and
        # zero_grad() clears just the remaining slices (everything autograd accumulates into, plus small direct ones)
        # with one vg_zero_segments launch: the 0.9 GB memset and the weight-gradient GEMMs' read of C disappear.
        self.direct: List[nn.Parameter] = []
        self._calibrating = False
        self._direct_calls: Dict[int, int] = {}
        self._autograd_hits: Dict[int, int] = {}
        self._overwrite: set = set()          # id(param): first direct write of a step uses beta = 0
        self._written: set = set()
        self._segments: Optional[List[Optional[Tuple[torch.Tensor, torch.Tensor, int]]]] = None
        if direct_wgrad:
            for grp in self.groups:
                for p in grp.params:
                    p._vg_main_grad = p.grad
                    p._vg_arena = self
                    self.direct.append(p)

    # ------------------------------------------------------------------ calibration of the overwrite set
    # elements.  Smaller slices are simply cleared (one launch clears all of them).  1.5 M keeps the [1024, 1024] output
    # projections on the accumulate path: their weight-gradient GEMMs run split-K (red.global.add into C), which
    # would need its own clear of C for beta = 0.
    _BIG = 3 << 19
    _HYPER_DEPTH = 4

    def calibrate_next(self) -> None:
        """observe, during the next step, which parameters receive their gradient only through wgrad_beta() sites.
        Tensor hooks see the gradient autograd is about to accumulate — None when an op wrote the slice itself (torch
        calls tensor and post-accumulate hooks for undefined gradients too, so the VALUE has to be looked at)."""
        self._calibrating = True
        self._direct_calls, self._autograd_hits = {}, {}
        self._overwrite, self._segments = set(), None

        def seen(g, i):
            if g is not None:
                self._autograd_hits[i] = self._autograd_hits.get(i, 0) + 1

        self._cal_handles = [p.register_hook(lambda g, i=id(p): seen(g, i)) for p in self.direct]

    def finish_calibration(self) -> None:
        self._calibrating = False
        for h in getattr(self, "_cal_handles", []):
            h.remove()
        self._cal_handles = []
        self._overwrite = {i for i, n in self._direct_calls.items() if n > 0 and self._autograd_hits.get(i, 0) == 0}
        segments = []
        for grp in self.groups:
            runs: List[List[int]] = []          # [offset, length] in elements, merged
            for k, (p, o) in enumerate(zip(grp.params, grp.offsets)):
                end = grp.offsets[k + 1] if k + 1 < len(grp.offsets) else grp.numel
                if id(p) in self._overwrite and p.numel() >= self._BIG:
                    continue
                if runs and runs[-1][0] + runs[-1][1] == o:
                    runs[-1][1] = end - runs[-1][0]
                else:
                    runs.append([o, end - o])
            total = sum(r[1] for r in runs)
            if not runs:
                segments.append((None, None, 0))
            elif total * 2 > grp.numel or len(runs) > 4096:
                segments.append(None)                                  # not worth it: whole-arena clear
            else:
                dev = grp.g.device
                segments.append((torch.tensor([r[0] for r in runs], dtype=torch.int64, device=dev),
                                 torch.tensor([r[1] for r in runs], dtype=torch.int64, device=dev), len(runs)))
        self._segments = segments

    # ------------------------------------------------------------------ per-step protocol
    def zero_grad(self) -> None:
        """clear what is accumulated into: the whole arena (0.9 GB at HBM speed ≈ 0.15 ms) before calibration, afterwards
        only the slices outside the overwrite set."""
        for gi, grp in enumerate(self.groups):
            seg = self._segments[gi] if self._segments is not None else None
            if seg is None:
                grp.g.zero_()
            elif seg[2] > 0:
                L.call("vg_zero_segments", L.ptr(grp.g), L.ptr(seg[0]), L.ptr(seg[1]), seg[2], L.stream())
        self._written = set()
        self.micro_batch = 0

    def wgrad_beta(self, param: Optional[nn.Parameter] = None) -> float:
        """beta of a direct gradient write into ``param``'s slice: 0 for the first write of a step into a parameter of
        the overwrite set (its slice was not cleared), 1 otherwise."""
        if param is None:
            return 1.0
        i = id(param)
        if self._calibrating:
            self._direct_calls[i] = self._direct_calls.get(i, 0) + 1
            return 1.0
        if i in self._overwrite and param.numel() >= self._BIG and i not in self._written:
            self._written.add(i)
            return 0.0
        return 1.0

    def end_micro_batch(self) -> None:
        self.micro_batch += 1

    def grad_ready(self, param: nn.Parameter) -> None:
        if self.on_grad_ready is not None:
            self.on_grad_ready(param)

    def adamw_step(self, lr: float, beta1: float = 0.9, beta2: float = 0.98, eps: float = 1e-8,
                   grad_scale: float = 1.0, use_device_hyper: bool = False) -> None:
        """torch.optim.AdamW semantics over both arenas; refreshes the bf16 shadows in the same pass."""
        bc1, bc2 = self.stage_hyper(lr, beta1, beta2)
        if use_device_hyper and not _capturing():
            self.upload_hyper()
        for grp, hyper in zip(self.groups, self.hyper):
            hp = L.ptr(hyper) if use_device_hyper else None
            L.call("vg_adamw_step", L.ptr(grp.p), L.ptr(grp.g), L.ptr(grp.m), L.ptr(grp.v),
                   L.ptr(grp.shadow) if grp.shadow is not None else None, grp.numel, lr, beta1, beta2, eps,
                   grp.weight_decay, bc1, bc2, grad_scale, hp, L.stream())

    def stage_hyper(self, lr: float, beta1: float = 0.9, beta2: float = 0.98):
        """advance the step counter and write {lr, 1/bc1, 1/sqrt(bc2), 1-lr*wd} to the pinned staging buffers."""
        self.step_count += 1
        self.generation += 1
        bc1 = 1.0 - beta1 ** self.step_count
        bc2 = 1.0 - beta2 ** self.step_count
        k = self.step_count % self._HYPER_DEPTH
        ev = self._hyper_events[k]
        if ev is not None and not _capturing():   # (a capture only advances the host counter; nothing reads the ring)
            ev.synchronize()                      # the upload that last read this ring entry has executed
        self._hyper_host = self._hyper_ring[k]
        for grp, host in zip(self.groups, self._hyper_host):
            host[0], host[1], host[2], host[3] = lr, 1.0 / bc1, bc2 ** -0.5, 1.0 - lr * grp.weight_decay
        return bc1, bc2

    def upload_hyper(self) -> None:
        """enqueue, on the current stream and OUTSIDE any captured graph, the copy of the staged scalars to the device
        buffers the AdamW kernels read (use_device_hyper): stream order puts it after the previous step's kernels and
        before this step's, whatever the host's lead over the GPU."""
        for hyper, host in zip(self.hyper, self._hyper_host):
            hyper.copy_(host, non_blocking=True)
        if torch.cuda.is_available() and self.hyper[0].is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._hyper_events[self.step_count % self._HYPER_DEPTH] = ev

    # ------------------------------------------------------------------ data-parallel buckets
    def buckets(self, bucket_bytes: int = 64 << 20) -> List[Tuple[torch.Tensor, List[nn.Parameter]]]:
        """contiguous gradient slices of ~bucket_bytes with the parameters they hold, in arena order."""
        out = []
        self.bucket_ranges: List[Tuple[int, int, int]] = []      # (group index, start, end) per bucket
        for gi, grp in enumerate(self.groups):
            start, members = 0, []
            for i, p in enumerate(grp.params):
                end = grp.offsets[i + 1] if i + 1 < len(grp.params) else grp.numel
                members.append(p)
                if (end - start) * 4 >= bucket_bytes or i + 1 == len(grp.params):
                    out.append((grp.g[start:end], members))
                    self.bucket_ranges.append((gi, start, end))
                    start, members = end, []
        return out

    def begin_step(self, lr: float, beta1: float = 0.9, beta2: float = 0.98, use_device_hyper: bool = False):
        """advance the step counter and stage this step's scalars; must precede the first adamw_bucket()."""
        self._bc = self.stage_hyper(lr, beta1, beta2)
        self._step_args = (lr, beta1, beta2, use_device_hyper)
        if use_device_hyper and not _capturing():
            self.upload_hyper()                   # (a captured step gets its scalars uploaded before each replay)

    def adamw_bucket(self, bucket_index: int, eps: float = 1e-8, grad_scale: float = 1.0) -> None:
        """AdamW (+ bf16 shadow refresh) of ONE bucket of buckets(): lets the optimizer run, bucket by bucket, as soon as
        a bucket's gradients are final (and all-reduced) — HBM-bound work hidden under the rest of backward."""
        gi, start, end = self.bucket_ranges[bucket_index]
        grp = self.groups[gi]
        lr, beta1, beta2, dev_hyper = self._step_args
        bc1, bc2 = self._bc
        sl = slice(start, end)
        L.call("vg_adamw_step", L.ptr(grp.p[sl]), L.ptr(grp.g[sl]), L.ptr(grp.m[sl]), L.ptr(grp.v[sl]),
               L.ptr(grp.shadow[sl]) if grp.shadow is not None else None, end - start, lr, beta1, beta2, eps,
               grp.weight_decay, bc1, bc2, grad_scale, L.ptr(self.hyper[gi]) if dev_hyper else None, L.stream())

    def total_numel(self) -> int:
        return sum(g.numel for g in self.groups)

    # ------------------------------------------------------------------ resume
    def optimizer_state_dict(self) -> Dict[str, object]:
        """AdamW state keyed by PARAMETER NAME (exp_avg / exp_avg_sq as tensors of the parameter's shape, like
        ``torch.optim.AdamW.state_dict()['state']`` entries) + the step count, so that it survives a change of arena
        layout (bucket order, alignment).  The parameters themselves are in ``model.state_dict()``."""
        state = {}
        for grp in self.groups:
            for name, p, o in zip(grp.names, grp.params, grp.offsets):
                n = p.numel()
                state[name] = {"exp_avg": grp.m[o:o + n].view(p.shape).detach().cpu().clone(),
                               "exp_avg_sq": grp.v[o:o + n].view(p.shape).detach().cpu().clone()}
        return {"step": self.step_count, "state": state}

    def snapshot(self) -> Dict[str, object]:
        """parameters, moments, shadows and the step count — TrainStep takes one before its calibration / warm-up / capture
        steps (real optimizer updates on the example batch) and restores it afterwards, so that building a TrainStep leaves
        a freshly initialised or just-resumed model exactly as it found it."""
        return {"step": self.step_count,
                "groups": [(g.p.clone(), g.m.clone(), g.v.clone(), g.shadow.clone() if g.shadow is not None else None)
                           for g in self.groups]}

    def restore(self, snap: Dict[str, object]) -> None:
        self.step_count = int(snap["step"])
        self.generation += 1
        for g, (p, m, v, sh) in zip(self.groups, snap["groups"]):
            g.p.copy_(p)
            g.m.copy_(m)
            g.v.copy_(v)
            if sh is not None:
                g.shadow.copy_(sh)

    def load_optimizer_state_dict(self, sd: Dict[str, object]) -> None:
        self.step_count = int(sd["step"])
        self.generation += 1
        for grp in self.groups:
            for name, p, o in zip(grp.names, grp.params, grp.offsets):
                n = p.numel()
                st = sd["state"][name]
                grp.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                grp.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))

    def refresh_shadow(self) -> None:
        """re-derive the bf16 weight shadows after parameters were loaded into the arena (model.load_state_dict)."""
        self.generation += 1
        for grp in self.groups:
            if grp.shadow is not None:
                L.call("vg_cast_f32_to_bf16", L.ptr(grp.p), L.ptr(grp.shadow), grp.numel, L.stream())


def cosine_lr(step: int, base_lr: float, min_lr: float, flat_steps: int, total_steps: int) -> float:
    """training_lib/optimizer.py:58-107 for the VAE-GSLM recipe: flat for `flat_steps`, then cosine to min_lr."""
    import math
    if step < flat_steps:
        return base_lr
    t, tmax = step - flat_steps, max(1, total_steps - flat_steps)
    return min_lr + (base_lr - min_lr) * (1 + math.cos(math.pi * min(t, tmax) / tmax)) / 2
