"""Data-parallel gradient all-reduce for the VAE-GSLM training step.

The reference trains with Lightning's ``ddp`` strategy (scripts/train.py:93-95): DistributedDataParallel,
25 MB buckets, gradient MEAN across ranks, fired on every micro-batch.  Here (one process per GPU, NCCL over
NVLink/NVSwitch through torch.distributed) the buckets are contiguous slices of the ParamArena gradient
buffer (no flatten copies); a bucket is all-reduced on a side stream as soon as backward has produced all
of its gradients, overlapping the rest of backward; and the collective only runs on the last micro-batch
of an accumulation window.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .arena import ParamArena


class GradReducer:
    def __init__(self, arena: ParamArena, bucket_bytes: int = 64 << 20, process_group=None,
                 overlap: bool = True, compress_bf16: bool = False) -> None:
        """``compress_bf16`` (opt-in, off by default = the reference's fp32 DDP reduction): each bucket is rounded to
        bf16 into a persistent staging buffer, all-reduced there and widened back — half the bytes on the wire, the
        trade torch's ``bf16_compress_hook`` makes."""
        self.arena = arena
        self.compress_bf16 = compress_bf16
        self._stage: Dict[int, torch.Tensor] = {}
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets = arena.buckets(bucket_bytes)
        self.overlap = overlap
        self._bucket_of: Dict[int, int] = {}
        self._pending: List[int] = []
        self._size: List[int] = []
        for bi, (_, members) in enumerate(self.buckets):
            self._size.append(len(members))
            for p in members:
                self._bucket_of[id(p)] = bi
        self._handles = []
        self.enabled = False           # armed only for the micro-batch that closes an accumulation window
        # optional callable(bucket_index) enqueued on the communication stream right after a bucket's all-reduce:
        # TrainStep installs the bucket-wise AdamW there, so the optimizer overlaps the rest of backward too
        self.after_bucket = None
        self.is_cuda = self.buckets[0][0].is_cuda
        self.comm_stream = torch.cuda.Stream() if self.is_cuda else None
        self._launched: List[bool] = []
        # streams other than the current one on which gradients of a bucket may still be in flight when the bucket's
        # last gradient is announced (LVTR runs the utterance encoder / diffusion decoder branch on a side stream)
        self.extra_streams: List["torch.cuda.Stream"] = []
        for _, members in self.buckets:
            for p in members:                 # autograd-accumulated gradients announce themselves through a hook;
                p.register_post_accumulate_grad_hook(self._hook)
        arena.on_grad_ready = self._ready     # gradients written directly by ops._wgrad/_bgrad/_RMSNorm call this

    # ------------------------------------------------------------------ per-backward protocol
    def calibrate_next(self) -> None:
        """count, during the next backward, how many times each parameter's gradient is announced (a parameter used
        twice in the graph is announced twice) and use those counts afterwards: a bucket must not be reduced / stepped
        before its LAST contribution.  The calibration backward launches nothing early."""
        self._calibrating = True
        self._seen: Dict[int, int] = {}

    def prepare(self, last_micro_batch: bool = True) -> None:
        # armed on the micro-batch that closes an accumulation window, when there is something to do per bucket:
        # an all-reduce (world > 1) and / or the bucket-wise optimizer step (after_bucket)
        self.enabled = last_micro_batch and (self.world > 1 or self.after_bucket is not None)
        self._pending = list(self._size)
        self._launched = [False] * len(self.buckets)
        self._handles = []

    def _hook(self, param) -> None:
        self._ready(param)

    def _ready(self, param) -> None:
        if not self.enabled:
            return
        bi = self._bucket_of.get(id(param))
        if bi is None:
            return
        if getattr(self, "_calibrating", False):
            self._seen[id(param)] = self._seen.get(id(param), 0) + 1
            return
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self.overlap:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        if self._launched[bi]:
            return
        self._launched[bi] = True
        flat = self.buckets[bi][0]
        if self.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            for st in self.extra_streams:
                self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                if self.world > 1 and self.compress_bf16:
                    stage = self._staging(bi, flat)
                    stage.copy_(flat)
                    dist.all_reduce(stage, op=dist.ReduceOp.AVG, group=self.group)
                    flat.copy_(stage)
                elif self.world > 1:
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
                if self.after_bucket is not None:
                    self.after_bucket(bi)
        elif self.world > 1:       # gloo (CPU tests): no AVG op
            if self.compress_bf16:
                stage = self._staging(bi, flat)
                stage.copy_(flat)
                dist.all_reduce(stage, op=dist.ReduceOp.SUM, group=self.group)
                flat.copy_(stage)
                h = None
            else:
                h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._handles.append((h, flat, bi))
        elif self.after_bucket is not None:
            self.after_bucket(bi)

    def _staging(self, bi: int, flat: torch.Tensor) -> torch.Tensor:
        if bi not in self._stage:
            self._stage[bi] = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
        return self._stage[bi]

    def finish(self) -> None:
        """reduce whatever has not been launched yet (parameters unused in this step never fire a hook) and make
        the compute stream wait for the collectives."""
        if getattr(self, "_calibrating", False):
            # (also when nothing was armed — world 1 without a bucket-wise optimizer: the flag must not outlive its step)
            self._calibrating = False
            if self.enabled:
                self._size = [sum(self._seen.get(id(p), 0) for p in members) for _, members in self.buckets]
        if not self.enabled:
            return
        for bi in range(len(self.buckets)):
            self._launch(bi)
        if self.is_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            for h, flat, bi in self._handles:
                if h is not None:
                    h.wait()
                flat.div_(self.world)
                if self.after_bucket is not None:
                    self.after_bucket(bi)
        self.enabled = False
