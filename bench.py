#!/usr/bin/env python
"""bench.py — VAE-GSLM hot-path benchmark (contract: see the task brief / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (N>1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K ...   the reference algorithm (oracle port) on host cores

A "step" is one full training step of BASELINE.json configs[1] on one synthetic micro-batch per GPU:
forward + backward of LVTR (bf16 activations, tcgen05 GEMMs), data-parallel gradient all-reduce (N>1) and the
fused AdamW update.  `value` = frames of ALL ranks / max-over-ranks device time, inputs resident in HBM;
`e2e` repeats the measurement with pinned HOST inputs copied in and the loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = os.path.join(ROOT, "vae_gslm_b200", "configs", "train", "speech", "vae-gslm.yaml")
N_MELS, VOCAB, KW = 80, 200, 0.04


def train_flops_per_frame(T: int) -> float:
    """fwd MAC/frame = LM-side 204,228,384 + attention 16,384·(T+1) + conv encoder 6,345,216 +
    UNet 14,527,488; train FLOPs = 6 × MAC."""
    return 6.0 * (204_228_384 + 16_384 * (T + 1) + 6_345_216 + 14_527_488)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "src": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "src": "fallback"}


def synthetic_batch(B, T, rank, pin=False):
    """SURVEY §8d synthetic inputs: tokens ~ randint(0,200), mel ~ N(0,1), utterance crop N(0,1)[B,150,80] with
    lengths ~ U{100..150}; full-length regime (all lengths = T)."""
    g = torch.Generator().manual_seed(1234 + rank)
    tokens = torch.randint(0, VOCAB, (B, T), generator=g)
    mel = torch.randn(B, T, N_MELS, generator=g)
    x = torch.cat([tokens[..., None].float(), mel], -1)
    ul = torch.randint(100, 151, (B,), generator=g)
    utt = torch.randn(B, 150, N_MELS, generator=g)
    batch = {"x": x, "mask": torch.ones(B, T, dtype=torch.bool), "utterance": utt,
             "utt_mask": torch.arange(150)[None, :] < ul[:, None]}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    rows.append((float(parts[0]), float(parts[1]), parts[3:7]))
                except ValueError:
                    pass
        os.unlink(self.path)
        if rows:
            sm = sorted(r[0] for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = rows[0][1]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[2][i].lower().startswith("active") for r in rows)]
            out["samples"] = len(rows)
        return out


# ====================================================================================== our arm
def run_ours(args):
    import torch.distributed as dist
    from vae_gslm_b200 import _lib, ops
    from vae_gslm_b200.arena import ParamArena
    from vae_gslm_b200.dp import GradReducer
    from vae_gslm_b200.hparams.hp import Hparams
    from vae_gslm_b200.models.speech.lvtr import LVTR
    from vae_gslm_b200.trainers.speech.lvtr import assemble_loss
    from vae_gslm_b200.training_lib.trainer import init_weights
    from vae_gslm_b200.utils.tensormask import TensorMask

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py measures the CUDA path: no GPU visible (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # (Tried and rejected, profiles/r01_scaling.md: capping NCCL to 8 CTAs and the persistent GEMM grids to the other
        #  140 SMs — NCCL_MAX_CTAS=8 VG_GEMM_SMS=140 — makes the all-reduce too slow to hide: 22.9 vs 21.2 ms at N=8.)
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    _lib.load()

    B, T = args.batch, args.frames
    torch.manual_seed(0)                                    # identical weights on every rank
    hp = Hparams.from_yamlfile(CFG)
    model = LVTR(hp.model, input_dim=N_MELS)
    model.apply(init_weights)
    model = model.to(dev).set_compute_dtype(torch.bfloat16)
    arena = ParamArena(model, weight_decay=hp.training.optimizer.weight_decay)
    # VG_DP_BF16=1: opt-in bf16 staging of the gradient all-reduce (dp.py); off = the reference's fp32 DDP reduction
    dp_bf16 = world > 1 and os.environ.get("VG_DP_BF16", "0") == "1"
    reducer = GradReducer(arena, bucket_bytes=64 << 20, compress_bf16=dp_bf16)
    lr = hp.training.optimizer.lr

    host = synthetic_batch(B, T, rank, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    from vae_gslm_b200.trainers.speech.lvtr import TrainStep
    train_step = TrainStep(model, arena, reducer, resident, lr=lr, kld_weight=KW,
                           use_cuda_graph=not args.no_cuda_graph)
    if train_step.graph is None and not args.no_cuda_graph:
        # a failed capture leaves torch's CUDA RNG in capture mode: restart this process in eager mode
        print("cuda-graph capture failed, re-running eagerly:", train_step.capture_error, file=sys.stderr, flush=True)
        os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-cuda-graph"])

    def step(batch, from_host):
        if from_host:
            train_step.load(batch)                          # pinned host → static device buffers (H2D every step)
        loss = train_step()
        if from_host:
            return float(loss.item())                       # device→host read of the step's result
        return loss

    def timed(nsteps, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(nsteps):
            last = step(host if from_host else resident, from_host)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), last

    for _ in range(args.warmup):
        step(resident, False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, last_loss = timed(args.steps, False)
    launches = train_step.launches_per_step * args.steps     # libvgslm kernels (graph replays re-launch all of them)
    clocks = sampler.stop()
    ms_e2e, loss_host = timed(args.steps, True)

    frames_per_step = B * T * world
    value = frames_per_step * args.steps / (ms / 1e3)
    e2e_value = frames_per_step * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (gemm_tc_kernel): one extra, instrumented step (not part of the timing)
    ops.PROFILE = []
    # eager and serial on purpose (no side streams): CUDA events around every GEMM launch, one kernel at a time
    mark = os.environ.get("VG_BENCH_MARK") == "1"          # bracket this step with spin_kernel markers for ncu launch lists
    if mark:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()                        # ncu --profile-from-start off: only this step is profiled
        torch.cuda._sleep(1000)
    # The eager pass is HOST-bound (~1300 launches at 10-20 us of Python each against ~22 ms of kernels): an event pair
    # around a short GEMM then also measures the idle gap before its launch arrives (the out-projection forward read
    # 42 us in this pass and 24 us launched back to back, profiles/r02_gemm_epilogue.md).  A ~30 ms spin kernel in front
    # lets the host run ahead, so that the device executes the pass back to back and the events bracket kernels only.
    if not mark:                                        # (under ncu every launch is serialised anyway)
        torch.cuda._sleep(int(0.03 * 1.9e9))
        ops.PROFILE_SPIN = int(60e-6 * 1.9e9)           # and 60 us in front of every timed GEMM keeps the device behind
    train_step._run_eager(device_hyper=False, serial=True)
    ops.PROFILE_SPIN = 0
    if mark:
        torch.cuda._sleep(1000)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    gemm_ms = sum(p_[0].elapsed_time(p_[1]) for p_ in prof)
    gemm_flops = sum(p_[2] for p_ in prof)
    if os.environ.get("VG_BENCH_GEMM_TABLE") == "1" and rank == 0:
        agg = {}
        for p_ in prof:
            a_ = agg.setdefault(p_[4], [0, 0.0, 0.0])
            a_[0] += 1
            a_[1] += p_[0].elapsed_time(p_[1])
            a_[2] += p_[2]
        print("| M | N | K | tA | tB | out | act | bias | launches | ms | TFLOP/s |\n|---|---|---|---|---|---|---|---|---:|---:|---:|", file=sys.stderr)
        for k_, (n_, ms_, fl_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("| " + " | ".join(str(x) for x in k_) + f" | {n_} | {ms_:.3f} | {fl_ / ms_ / 1e9:.0f} |", file=sys.stderr)
    peaks = measured_peaks()
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    step_ms = ms / args.steps
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM, all instances of one step)",
                "achieved": round(achieved, 1), "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["bf16_sustained"], 4), "peak_source": peaks["src"] + " sustained",
                # the GEMMs are timed one at a time (eager, serial pass): against the BURST peak the same number reads
                "peak_burst": peaks["bf16_burst"], "frac_vs_burst": round(achieved / peaks["bf16_burst"], 4),
                # DRAM read+write bytes of ONE representative launch (FFN1 forward, bias+GELU+saved derivative, 8000x4096x1024)
                # from `ncu --set full`: profiles/r02_gemm_ffn1_ncu.md (algorithmic bytes of that launch: 155.8e6)
                "traffic": 112.9e6, "traffic_launch": "ffn1 fwd 8000x4096x1024 (ncu --set full of the final build: profiles/r02_gemm_ffn1_ncu.md)",
                "launches_per_step": len(prof), "gemm_ms_per_step": round(gemm_ms, 3),
                "share_of_step": round(gemm_ms / step_ms, 3)}

    result = {
        "metric": "train_mel_frames_per_sec", "value": round(value, 1), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"VAE-GSLM bf16 training step (fwd+bwd+allreduce+AdamW), {T / 50:.0f} s segments",
                   "per_gpu_batch": B, "frames_per_seq": T, "global_batch": B * world, "parallelism": f"dp{world}" + ("+bf16-grad-allreduce" if dp_bf16 else ""),
                   "params": 226_957_564, "l2": "working set (454 MB bf16 weights + activations) exceeds the 126 MB L2",
                   "init": "random (reference init rules)"},
        "per_gpu_frames_per_sec": round(value / world, 1),
        "model_tflops_per_gpu": round(train_flops_per_frame(T) * value / world / 1e12, 1),
        "mfu_vs_measured_sustained": round(train_flops_per_frame(T) * value / world / 1e12 / peaks["bf16_sustained"], 4),
        "mfu_vs_nominal_2250": round(train_flops_per_frame(T) * value / world / 1e12 / 2250.0, 4),
        "e2e": {"value": round(e2e_value, 1), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": loss_host},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "loss": float(last_loss), "step_mode": train_step.mode,
    }
    # ---- the other BASELINE shapes on the same model / arena / reducer (new static buffers, new capture): the repo
    #      config's shape (8 x 640 frames, configs/train/speech/vae-gslm.yaml:173-184) and configs[4] (60 s = 3000 frames)
    if not args.no_shapes:
        result["shapes"] = {}
        for (b2, t2) in ((8, 640), (2, 3000)):
            train_step.graph = None
            res2 = {k: v.to(dev) for k, v in synthetic_batch(b2, t2, rank).items()}
            ts2 = TrainStep(model, arena, reducer, res2, lr=lr, kld_weight=KW, use_cuda_graph=not args.no_cuda_graph)
            if ts2.graph is None and not args.no_cuda_graph:
                # (a failed capture leaves torch's CUDA RNG in capture mode: nothing more can run in this process)
                result["shapes"][f"B{b2}_T{t2}"] = {"unavailable": "cuda-graph capture failed: " + str(ts2.capture_error)[:200]}
                args.no_decode = args.no_gpu_reference = args.no_cpu_baseline = True
                break
            for _ in range(3):
                ts2()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n2 = max(5, args.steps // 2)
            e0.record()
            for _ in range(n2):
                ts2()
            e1.record()
            torch.cuda.synchronize()
            ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
            v2 = b2 * t2 * world * n2 / (float(ms2) / 1e3)
            result["shapes"][f"B{b2}_T{t2}"] = {
                "value": round(v2, 1), "unit": "frames/s", "ms_per_step": round(float(ms2) / n2, 3), "steps": n2,
                "per_gpu_frames_per_sec": round(v2 / world, 1), "step_mode": ts2.mode,
                "model_tflops_per_gpu": round(train_flops_per_frame(t2) * v2 / world / 1e12, 1),
                "mfu_vs_measured_sustained": round(train_flops_per_frame(t2) * v2 / world / 1e12 / peaks["bf16_sustained"], 4),
                "mfu_vs_nominal_2250": round(train_flops_per_frame(t2) * v2 / world / 1e12 / 2250.0, 4)}
            ts2.graph = None
            del ts2, res2
            torch.cuda.empty_cache()
    if world > 1:
        # data-parallel invariant: after the same number of steps every rank holds bit-identical parameters
        chk = torch.stack([torch.stack([g.p.double().sum(), (g.p.double() ** 2).sum()]) for g in arena.groups]).reshape(-1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        result["dp_check"] = {"params_identical_across_ranks": bool(all(torch.equal(allc[0], c) for c in allc)),
                              "checksum": [float(x) for x in allc[0]], "optimizer_steps": arena.step_count}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline_sample(lr)
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        train_step.graph = None
        torch.cuda.empty_cache()
        result["gpu_reference"] = gpu_reference_block(dev, lr, B, T)
        if "frames_per_sec" in result["gpu_reference"]:
            result["gpu_reference"]["ours_over_reference"] = round(value / result["gpu_reference"]["frames_per_sec"], 2)
    if rank == 0 and world == 1 and not args.no_decode:
        result["decode"] = decode_bench(model, dev, peaks, batches=(1, 2, 4, 8, 16, 32, 64, 128, 256))
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        # Tear-down.  Graphs that captured NCCL kernels must be destroyed before the communicator: release them, drain the
        # device, meet at a barrier, then destroy the process group — on a watchdog, because destroy_process_group() was
        # observed (round 1) to hang behind captured NCCL nodes; if it has not returned after 30 s the process leaves
        # without the collective destructor (the result line above is already flushed).
        import threading
        train_step.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()

        threading.Thread(target=_destroy, daemon=True).start()
        if not done.wait(30.0):
            print(f"rank {rank}: destroy_process_group() still blocked after 30 s, exiting without it", file=sys.stderr, flush=True)
            os._exit(0)


def decode_bench(model, dev, peaks, batches=(1, 64, 256), prompt=150, gen=500, ddim=True):
    """cached generation (configs[2]: 3 s prompt → 10 s continuation): prefill of prompt + BOS, then `gen` - 1 single-token
    steps replayed from one CUDA graph (the cache position lives on the device), i.e. the cache grows 151 → 650 and the
    mean number of attended keys is ~400.  frames/s against the HBM roofline of SURVEY §8d summed over exactly these steps:
    bytes(B, Tk) = 408.7 MB of weights + B·65,536·(Tk + 1) of KV cache per step."""
    from vae_gslm_b200.trainers.speech.sampler import GraphedStep
    out = {}
    model.eval()
    for B in batches:
        g = torch.Generator().manual_seed(7)
        prior = torch.cat([torch.randint(0, VOCAB, (B, prompt, 1), generator=g).float(),
                           torch.randn(B, prompt, 4, generator=g)], -1).to(dev)
        model.transformer[0].cache_len_hint = prompt + 1 + gen + 8
        o = model.step(prior, past_kv=None, temperature=0.85, token_temperature=0.85, push_init_state=True)
        state, kv = o["output"][:, -1:], o["kv"]
        warm = 3
        for _ in range(warm):
            o = model.step(state, past_kv=kv, temperature=0.85, token_temperature=0.85)
            state, kv = o["output"], o["kv"]
        graphed = GraphedStep(model, state, kv, temperature=0.85, token_temperature=0.85)
        graphed()
        torch.cuda.synchronize()
        tk0 = kv[0].cache.length                   # keys cached before the first timed step
        steps = gen - 1 - warm - 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graphed()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        assert kv[0].cache.length == tk0 + steps
        mean_tk = tk0 + (steps - 1) / 2.0          # step i attends tk0 + i cached keys + its own
        bytes_step = 408.7e6 + B * 65536 * (mean_tk + 1)
        roof = B * peaks["hbm_gbs"] * 1e9 / bytes_step
        eng = model.__dict__.get("_decode_engines", {}).get(B)
        path = type(eng[1]).__name__ if (eng and model.use_decode_engine) else "layer-by-layer (tcgen05 GEMM; <= 1024-feature layers on vg_skinny_linear)"
        out[f"B{B}"] = {"frames_per_sec": round(B / (ms / 1e3), 1), "ms_per_step": round(ms, 4),
                        "hbm_roofline_frames_per_sec": round(roof, 1), "frac": round(B / (ms / 1e3) / roof, 4),
                        "achieved_gbs": round(bytes_step / (ms / 1e3) / 1e9, 1), "steps": steps,
                        "tk_first": tk0, "tk_last": tk0 + steps - 1, "mean_tk": round(mean_tk, 1),
                        "mode": "single-token step replayed from a CUDA graph", "path": path}
        del graphed, kv, o
        model.__dict__.pop("_decode_engines", None)          # one engine (packed weight stream) at a time
        torch.cuda.empty_cache()
    if ddim:
        out["ddim_decode"] = ddim_bench(model, dev)
    return out


def ddim_bench(model, dev, B=16, T=650, steps=100, eta=0.5):
    """diffusion decoding of generated frames (configs/infer: 100 DDIM steps, eta 0.5, 3 s + 10 s = 650
    frames), whole loop replayed from one CUDA graph; mel frames per second."""
    from vae_gslm_b200.trainers.speech.sampler import GraphedDecode
    g = torch.Generator().manual_seed(11)
    frames = torch.cat([torch.randint(0, VOCAB, (B, T, 1), generator=g).float(), torch.randn(B, T, 4, generator=g)], -1).to(dev)
    mask = torch.ones(B, T, dtype=torch.bool, device=dev)
    u_c = torch.randn(B, 128, generator=g).to(dev)
    dec = model.decoder
    saved = (dec.sampling_timesteps, dec.ddim_sampling_eta)
    dec.sampling_timesteps, dec.ddim_sampling_eta = steps, eta
    try:
        graphed = GraphedDecode(model, frames, mask, u_c)
        graphed()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            graphed()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
    finally:
        dec.sampling_timesteps, dec.ddim_sampling_eta = saved
    return {"mel_frames_per_sec": round(B * T / (ms / 1e3), 1), "ms": round(ms, 2), "batch": B, "frames": T,
            "ddim_steps": steps, "mode": "whole DDIM loop replayed from one CUDA graph"}


# ====================================================================================== reference arms
# The UNMODIFIED reference modules (vendored by __graft_entry__.build() into the git-ignored baseline/_ref, which travels to
# the GPU box) when present; otherwise the oracle port (oracle/lvtr_oracle.py — same arithmetic, proven against the real
# reference by tests/test_oracle_golden.py).  Used for: `--impl reference` (host cores), `cpu_baseline` (host cores, bounded
# sample) and `gpu_reference` (the same modules under torch.autocast(bf16) on the same B200 — the stock-PyTorch bar).
def _reference_modules():
    try:
        from baseline import vendor_reference as V
        if V.available():
            return V.load()
    except Exception as e:                                   # noqa: BLE001
        print("reference modules unavailable:", repr(e), file=sys.stderr)
    return None


class _ReferenceStep:
    """one training step (fwd + loss assembly of trainers/speech/lvtr.py:103-131 + bwd + AdamW) of the reference model"""

    def __init__(self, device, T, lr, autocast):
        mods = _reference_modules()
        self.device, self.autocast = torch.device(device), autocast
        torch.manual_seed(0)
        if mods is not None:
            LVTR, Hparams, TensorMask, masked_loss, cfg = mods
            hp = Hparams.from_yamlfile(cfg)
            if T > hp.model.transformer.rpe.maxpos:           # the reference's dense ALiBi buffer (position/alibi.py:8-17)
                hp.model.transformer.rpe.maxpos = (T + 63) // 64 * 64
            self.model = LVTR(hp.model, input_dim=N_MELS)

            def init(m):                                      # BaseTrainer.init_weights (training_lib/trainer.py:113-125)
                if getattr(m, "bias", None) is not None:
                    m.bias.data.zero_()
                if callable(getattr(m, "custom_weight_init", None)):
                    m.custom_weight_init(1.0)
            self.model.apply(init)
            self.model.to(self.device)
            self.TM, self.masked_loss, self.kind = TensorMask, masked_loss, "reference"
            params = list(self.model.parameters())
        else:
            from vae_gslm_b200.hparams.hp import Hparams
            hp = Hparams.from_yamlfile(CFG)
            self.sd = {k: v.to(self.device) for k, v in _oracle_state(None, hp).items()}
            self.sd = {k: v.detach().requires_grad_(v.requires_grad) for k, v in self.sd.items()}
            self.cfg, self.kind = hp.model.to_dict(), "port"
            params = [v for v in self.sd.values() if v.requires_grad]
        self.opt = torch.optim.AdamW([{"params": [p for p in params if p.ndim != 1]},
                                      {"params": [p for p in params if p.ndim == 1], "weight_decay": 0.0}],
                                     lr=lr, betas=(0.9, 0.98), weight_decay=0.1,
                                     **({"fused": True} if self.device.type == "cuda" else {}))

    def __call__(self, batch, rng=None):
        ctx = torch.autocast(self.device.type, dtype=torch.bfloat16) if self.autocast else torch.autocast(self.device.type, enabled=False)
        with ctx:
            if self.kind == "reference":
                out = self.model(self.TM(batch["x"], batch["mask"]), utterance=self.TM(batch["utterance"], batch["utt_mask"]))
                kld = self.masked_loss(out["log_q"] * 1.0, out["log_p"], fn=lambda a, b: (a - b))
                loss = out["decoder_output"] * 1.0 + kld * KW + out["ce_loss"] * 0.5 * KW
            else:
                from oracle import lvtr_oracle as O
                out = O.lvtr_forward(self.sd, self.cfg, batch["x"], batch["mask"], batch["utterance"], batch["utt_mask"], rng)
                loss = O.total_loss(out, KW)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss.detach()


def _oracle_state(model_or_none, hp):
    from vae_gslm_b200.models.speech.lvtr import LVTR
    from vae_gslm_b200.training_lib.trainer import init_weights
    if model_or_none is None:
        torch.manual_seed(0)
        model_or_none = LVTR(hp.model, input_dim=N_MELS)
        model_or_none.apply(init_weights)
    names = {n for n, _ in model_or_none.named_parameters()}
    return {k: v.detach().float().cpu().clone().requires_grad_(k in names) for k, v in model_or_none.state_dict().items()}


def _rng(B, T, seed=4321):
    g = torch.Generator().manual_seed(seed)
    return {"eps_q": torch.randn(B, T, 4, generator=g), "init_state": torch.rand(B, 1, 64, generator=g) * 2 - 1,
            "eps_p": torch.randn(B, T, 4, generator=g), "diff_t": torch.randint(0, 1000, (B,), generator=g),
            "diff_noise": torch.randn(B, T, N_MELS, generator=g)}


def cpu_baseline_sample(lr, B=2, T=1000):
    """the reference on the box's host cores (all of them), fp32: one warm-up step and one timed step of a bounded sample
    (B=2 of the 8 sequences of the workload)."""
    torch.set_num_threads(os.cpu_count() or 1)
    step = _ReferenceStep("cpu", T, lr, autocast=False)
    batch, rng = synthetic_batch(B, T, 0), _rng(B, T)
    step(batch, rng)                                          # warm-up at the timed shape (allocator, thread pools, oneDNN)
    t0 = time.perf_counter()
    step(batch, rng)
    dt = time.perf_counter() - t0
    return {"value": round(B * T / dt, 1), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": step.kind,
            "sample": f"1 warm-up + 1 timed fwd+bwd+AdamW step at B={B}, T={T} (fp32, torch CPU): {dt:.1f} s",
            "host_cpus": os.cpu_count()}


def gpu_reference_block(dev, lr, B, T, steps=4, warmup=3):
    """the stock-PyTorch bar: the reference's own modules — dense-mask SDPA, cuBLAS, cuDNN,
    ~40 small kernels per layer — under torch.autocast(bf16) with TF32 matmuls (scripts/train.py:44) and fused AdamW, on
    the SAME B200 and the same synthetic batch, timed with CUDA events.  Its RNG draws are its own (the timing does not
    depend on them)."""
    saved = torch.get_float32_matmul_precision()
    try:
        torch.set_float32_matmul_precision("high")
        step = _ReferenceStep(dev, T, lr, autocast=True)
        batch = {k: v.to(dev) for k, v in synthetic_batch(B, T, 0).items()}
        rng = {k: v.to(dev) for k, v in _rng(B, T).items()}
        for _ in range(warmup):
            step(batch, rng)
        torch.cuda.synchronize()
        # eager PyTorch is partly host-bound (~40 short kernels per layer): take the BEST of three timed blocks, so that a
        # busy host core does not flatter the comparison
        ms = float("inf")
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step(batch, rng)
            e1.record()
            torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1) / steps)
        out = {"frames_per_sec": round(B * T / (ms / 1e3), 1), "ms_per_step": round(ms, 2), "kind": step.kind,
               "precision": "torch.autocast(bf16) + TF32 matmul, fused AdamW", "batch": B, "frames": T, "steps": steps, "timing": "best of 3 blocks",
               "loss": float(loss), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    except Exception as e:                                   # noqa: BLE001
        out = {"unavailable": repr(e)[:300]}
    finally:
        torch.set_float32_matmul_precision(saved)
    del step
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vae_gslm_b200.hparams.hp import Hparams
    torch.set_num_threads(os.cpu_count() or 1)
    hp = Hparams.from_yamlfile(CFG)
    B, T = 1, args.frames                                  # bounded sample of the workload per step
    step = _ReferenceStep("cpu", T, hp.training.optimizer.lr, autocast=False)
    batch, rng = synthetic_batch(B, T, 0), _rng(B, T)
    for _ in range(args.warmup):
        step(batch, rng)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = float(step(batch, rng))
    dt = time.perf_counter() - t0
    value = B * T * args.steps / dt
    what = ("the UNMODIFIED reference modules (baseline/_ref: models/speech/lvtr.py:143-225, trainers/speech/lvtr.py:103-131)"
            if step.kind == "reference" else "oracle port of the PyTorch modules")
    print(json.dumps({
        "impl": "reference", "metric": "train_mel_frames_per_sec", "value": round(value, 1), "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"VAE-GSLM training step (fwd+bwd+AdamW), {T / 50:.0f} s segments — {what} on host cores",
                   "per_step_sample": f"B={B}, T={T}", "frames_per_seq": T, "parallelism": "cpu"},
        "cpu_baseline": {"value": round(value, 1), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": step.kind,
                         "sample": f"{args.steps} steps of B={B}, T={T} (fwd+bwd+AdamW), fp32"},
        "e2e": {"value": round(value, 1), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": loss}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="sequences per GPU (repo config: 8)")
    ap.add_argument("--frames", type=int, default=1000, help="frames per sequence (20 s segments at 50 Hz)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-shapes", action="store_true", help="skip the B=8,T=640 and B=2,T=3000 steps")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the stock-PyTorch bf16 step on the same GPU")
    ap.add_argument("--no-cuda-graph", action="store_true", help="run the training step eagerly (launch-bound)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
