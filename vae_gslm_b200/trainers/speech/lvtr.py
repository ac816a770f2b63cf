"""Loss assembly of the VAE-GSLM training step (reference ``trainers/speech/lvtr.py:103-145``) and the
single-process training step around it (``TrainStep``).  The Lightning shell of the reference is out of scope
(SURVEY §2 row 15); data-parallel plumbing lives in ``vae_gslm_b200.dp``.
"""
from __future__ import annotations

import os

from typing import Dict, Mapping, Optional

import torch

from ...training_lib.losses import masked_loss
from ...utils.tensormask import TensorMask


def kld_weight_at(global_step: int, kld_scale: float, warmup_kld: int = 0, zero_kld: int = 0) -> float:
    """KL weight schedule (:104-110): linear warm-up over `warmup_kld` steps after `zero_kld` zero steps.
    NB: with zero_kld = 0 and warmup_kld > 0 the weight is exactly 0 at global_step 0."""
    w = kld_scale
    if warmup_kld > 0 and zero_kld < (global_step + 1) <= warmup_kld:
        w = kld_scale * (global_step - zero_kld) / warmup_kld
    if zero_kld > 0 and global_step <= zero_kld:
        w = 0.0
    return w


def assemble_loss(output: Mapping, kld_weight, rec_loss_scale: float = 1.0, entropy_weight: float = 1.0,
                  token_kld_weight: float = 0.5, use_fused_kl: bool = True) -> Mapping[str, torch.Tensor]:
    """loss = rec·scale + kld·kw + ce·token_kld_weight·kw — every term a SUM over valid frames (:122-130).
    ``kld_weight`` may be a python float or a 0-dim device tensor (CUDA-graph replay)."""
    if use_fused_kl and entropy_weight == 1.0 and "kl_sum" in output:
        kld = output["kl_sum"]                       # computed inside the latent_back kernel
    else:
        kld = masked_loss(output["log_q"] * entropy_weight, output["log_p"], fn=lambda a, b: a - b)
    rec = output["decoder_output"]
    loss = rec * rec_loss_scale + kld * kld_weight
    ce = output.get("ce_loss")
    if ce is not None:
        loss = loss + ce * token_kld_weight * kld_weight
    return {"loss": loss, "kld": kld, "rec_loss": rec, "token_kld": ce, "kld_weight": kld_weight,
            "length": output["log_p"].mask.sum()}


def save_checkpoint(model, hp, directory: str, name: str = "last-cpt.ckpt", arena=None, global_step: Optional[int] = None) -> None:
    """the reference's COMPACT checkpoint (:294-296; training_lib/callbacks.py): ``model.state_dict()`` under ``name`` and
    the configuration as ``hp.yaml`` in ``directory`` — what ``inference/inferer.py:17-28`` and ``load_checkpoint`` read.
    With ``arena`` the AdamW moments and step go to ``<name>.optim`` next to it, for resuming."""
    import os
    os.makedirs(directory, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, os.path.join(directory, name))
    hp.save(os.path.join(directory, "hp.yaml"))
    if arena is not None:
        extra = arena.optimizer_state_dict()
        extra["global_step"] = global_step
        torch.save(extra, os.path.join(directory, name + ".optim"))


def load_checkpoint(model, directory: str, name: str = "last-cpt.ckpt", arena=None) -> Optional[int]:
    """load a compact checkpoint (ours or the reference's: same key names) into ``model``; with ``arena`` also the
    optimizer state if ``<name>.optim`` exists.  Returns the stored global step (None for a reference checkpoint)."""
    import os
    model.load_state_dict(torch.load(os.path.join(directory, name), map_location="cpu", weights_only=True), strict=False)
    step = None
    if arena is not None:
        if arena.groups[0].shadow is not None:
            arena.refresh_shadow()
        opt = os.path.join(directory, name + ".optim")
        if os.path.exists(opt):
            extra = torch.load(opt, map_location="cpu", weights_only=True)
            arena.load_optimizer_state_dict(extra)
            step = extra.get("global_step")
    return step


def make_model_input(tokens: TensorMask, mel: TensorMask) -> TensorMask:
    """channel-interleaved model input [B,T,1+n_mels]: token id as float ⊕ mel (:117-118)."""
    return TensorMask(tokens.value.to(mel.value.dtype), tokens.mask).expand().cat(mel)


class TrainStep:
    """zero-grad → [forward → loss → backward] x ``accumulate`` micro-batches → gradient all-reduce → fused AdamW, on
    static input buffers.

    ``accumulate`` is the recipe's ``training.gradient_accumulation`` (2 in configs/train/speech/vae-gslm.yaml): as in the
    reference's manual optimisation (trainers/speech/lvtr.py:147-158; training_lib/trainer.py:17-20) the UNSCALED losses of
    the micro-batches are back-propagated into the same gradients and the optimizer / LR schedule / KL schedule advance once
    per window; the data-parallel all-reduce runs on the last micro-batch only.
    With ``use_cuda_graph`` the whole sequence (≈5,000 kernel launches per micro-batch: ours, cuDNN for the utterance
    encoder, NCCL) is captured ONCE and replayed, which removes the Python/launch overhead that otherwise bounds the step.
    Per-step scalars that change (learning rate, Adam bias corrections, KL weight) live in device memory; they are staged
    in rotating pinned host buffers and uploaded by ordinary stream-ordered copies issued before each replay (never from
    inside the graph: a copy node would read the host buffer when it EXECUTES, by which time a host running ahead may
    have staged the next step's values).
    Building a TrainStep runs a calibration step, warm-up steps and the capture on the example batch; the arena is
    snapshotted before and restored after, so a fresh or just-resumed model is left exactly as it was found.
    """

    _streams: Dict[tuple, "torch.cuda.Stream"] = {}

    def __init__(self, model, arena, reducer, example_batch: Dict[str, torch.Tensor], *, lr: float,
                 kld_weight: float = 0.04, betas=(0.9, 0.98), eps: float = 1e-8, use_cuda_graph: bool = True,
                 warmup_iters: int = 3, overlap_grads: bool = True, accumulate: int = 1,
                 preserve_state: bool = True) -> None:
        self.pdl = int(os.environ.get("VG_TRAIN_PDL", "1"))
        self.model, self.arena, self.reducer = model, arena, reducer
        self.lr, self.betas, self.eps = lr, betas, eps
        self.accumulate = int(accumulate)
        assert self.accumulate >= 1
        # device-resident input buffers, one set per micro-batch of the accumulation window
        self.statics = [{k: v.clone() for k, v in example_batch.items()} for _ in range(self.accumulate)]
        self.static = self.statics[0]
        dev = next(iter(self.static.values())).device
        self.kw_dev = torch.full((), float(kld_weight), device=dev)
        self._kw_ring = [torch.full((), float(kld_weight)).pin_memory() if dev.type == "cuda" else torch.full((), float(kld_weight))
                         for _ in range(4)]
        self._kw_events = [None] * 4
        self._kw_i = 0
        self.loss = torch.zeros((), device=dev)
        self.terms = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.grad_stream = None
        self.overlap_grads = overlap_grads
        self.overlap_optimizer = overlap_grads          # bucket-wise AdamW behind each bucket's all-reduce
        self.mode = "eager"
        self.capture_error: Optional[str] = None
        # Every execution of the step body — warm-up, capture, eager — runs on ONE private stream.  autograd's
        # AccumulateGrad nodes remember the stream they were first used on (our parameters keep them alive through
        # the pre-assigned arena .grad views and the reducer hooks); if that were the default stream, capture on
        # another stream would be invalidated.  For the same reason every TrainStep of a process (bench.py builds one per
        # input shape on the same model) shares the stream of the first one.
        key = (dev.type, dev.index)
        if key not in TrainStep._streams:
            TrainStep._streams[key] = torch.cuda.Stream(device=dev)
        self.stream = TrainStep._streams[key]
        from ... import _lib
        snap = arena.snapshot() if (preserve_state and hasattr(arena, "snapshot")) else None
        n0 = _lib.launch_count
        if hasattr(self.reducer, "calibrate_next"):
            self.reducer.calibrate_next()                   # first step: learn how often each gradient is announced
        if hasattr(self.arena, "calibrate_next"):
            self.arena.calibrate_next()                     # … and which gradients are only ever written directly
        acc, self.accumulate = self.accumulate, 1           # calibration counts announcements of ONE micro-batch
        self._run_eager(device_hyper=False)                 # also the first warm-up iteration
        self.accumulate = acc
        if hasattr(self.arena, "finish_calibration"):
            self.arena.finish_calibration()
        self.launches_per_step = (_lib.launch_count - n0) * self.accumulate   # libvgslm kernels per step (gpu_launches)
        if use_cuda_graph:
            try:
                self._capture(warmup_iters)
                self.mode = "cuda-graph"
            except Exception as e:      # NB: a failed capture leaves torch's RNG in capture mode; callers restart
                self.graph = None
                self.capture_error = repr(e)
        if snap is not None:
            torch.cuda.synchronize(dev) if dev.type == "cuda" else None
            arena.restore(snap)

    def _run_eager(self, device_hyper: bool, serial: bool = False) -> None:
        """one eager step; ``serial`` switches every stream overlap off (bench.py's per-kernel timing pass)."""
        saved = (self.overlap_grads, self.overlap_optimizer, self.grad_stream, self.model.overlap_decoder)
        if serial:
            self.overlap_grads = self.overlap_optimizer = False
            self.grad_stream, self.model.overlap_decoder = None, False
            self.reducer.after_bucket = None
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self._body(device_hyper)
        torch.cuda.current_stream().wait_stream(self.stream)
        if serial:
            self.overlap_grads, self.overlap_optimizer, self.grad_stream, self.model.overlap_decoder = saved

    # the work of one step, on whatever stream is current
    def _body(self, device_hyper: bool) -> None:
        from ... import ops
        # VG_TRAIN_PDL (default 1; 0 = plain stream order): the tcgen05 GEMMs and RMSNorm forwards of the step are launched with the programmatic-dependent-
        # launch attribute (their barrier / TMEM set-up overlaps the tail of the kernel in front; no weight prefetch:
        # bit 0 only — in training the B operand of a GEMM is not always a static weight)
        with ops.pdl_mode(self.pdl):
            self._body_impl(device_hyper)

    def _body_impl(self, device_hyper: bool) -> None:
        from ... import ops
        if self.grad_stream is None and self.loss.is_cuda and self.overlap_grads:
            # one parameter-gradient stream per process and device: the reducer keeps every side stream it has seen in
            # `extra_streams` and waits for all of them — a stream left over from an earlier TrainStep would be waited
            # for from inside a later capture and invalidate it
            key = ("grad", self.loss.device.index)
            if key not in TrainStep._streams:
                TrainStep._streams[key] = torch.cuda.Stream(device=self.loss.device)
            self.grad_stream = TrainStep._streams[key]
        for side in (self.model.__dict__.get("_side_stream"), self.grad_stream):
            if side is not None and side not in self.reducer.extra_streams:
                self.reducer.extra_streams.append(side)
        ops.GRAD_STREAM = self.grad_stream
        self.arena.zero_grad()
        if self.overlap_optimizer:
            # bucket-wise AdamW on the communication stream, right behind each bucket's all-reduce
            self.arena.begin_step(self.lr, self.betas[0], self.betas[1], use_device_hyper=device_hyper)
            self.reducer.after_bucket = lambda bi: self.arena.adamw_bucket(bi, eps=self.eps)
        total = None
        for mb in range(self.accumulate):
            s = self.statics[mb]
            last = mb == self.accumulate - 1
            self.reducer.prepare(last_micro_batch=last)       # the collective (and the bucket-wise AdamW) only on the last
            out = self.model(TensorMask(s["x"], s["mask"]), utterance=TensorMask(s["utterance"], s["utt_mask"]))
            terms = assemble_loss(out, kld_weight=self.kw_dev)
            terms["loss"].backward()                          # unscaled, as the reference's manual_backward(loss)
            total = terms["loss"].detach() if total is None else total + terms["loss"].detach()
            if not last:
                self.arena.end_micro_batch()
                if self.grad_stream is not None:
                    torch.cuda.current_stream().wait_stream(self.grad_stream)
        ops.GRAD_STREAM = None
        if self.grad_stream is not None:
            torch.cuda.current_stream().wait_stream(self.grad_stream)     # parameter gradients are complete
        self.reducer.finish()
        if not self.overlap_optimizer:
            self.arena.adamw_step(self.lr, self.betas[0], self.betas[1], self.eps, use_device_hyper=device_hyper)
        self.loss.copy_(total)

    def _capture(self, warmup_iters: int) -> None:
        for _ in range(warmup_iters):                      # warm-up off the default stream, as torch requires
            self._run_eager(device_hyper=True)
        torch.cuda.synchronize()
        steps_before = self.arena.step_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._body(device_hyper=True)
        self.arena.step_count = steps_before               # capture advanced the host counter without running

    def load(self, batch: Dict[str, torch.Tensor], micro_batch: int = 0) -> None:
        """copy a (pinned host or device) batch into the static input buffers of one micro-batch.  If the batch carries an
        ``event`` attribute (data.dataset.PinnedBatch) an event recorded behind the copies is stored there: the assembler
        waits for it before it refills that pinned slot."""
        dst = self.statics[micro_batch]
        for k, v in batch.items():
            dst[k].copy_(v, non_blocking=True)
        if hasattr(batch, "event") and self.loss.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            batch.event = ev

    def set_kld_weight(self, kld_weight: float) -> None:
        """stage this step's KL weight (kld_weight_at(global_step, …), trainers/speech/lvtr.py:104-110) and enqueue its upload"""
        k = self._kw_i = (self._kw_i + 1) % len(self._kw_ring)
        if self._kw_events[k] is not None:
            self._kw_events[k].synchronize()
        self._kw_ring[k].fill_(float(kld_weight))
        self.kw_dev.copy_(self._kw_ring[k], non_blocking=True)
        if self.loss.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._kw_events[k] = ev

    def __call__(self, lr: Optional[float] = None, kld_weight: Optional[float] = None) -> torch.Tensor:
        if lr is not None:
            self.lr = lr
        if kld_weight is not None:
            self.set_kld_weight(kld_weight)
        if self.graph is not None:
            self.arena.stage_hyper(self.lr, self.betas[0], self.betas[1])     # rotating pinned staging …
            self.arena.upload_hyper()                                         # … uploaded in stream order, before the replay
            self.graph.replay()
        else:
            self._run_eager(device_hyper=False)
        return self.loss
