"""Data-parallel host logic on CPU: ParamArena layout + GradReducer (bucketed mean all-reduce) with world_size 2
over gloo.  The GPU path uses the same classes with NCCL on a side stream."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(golden):
    from vae_gslm_b200.arena import ParamArena
    from vae_gslm_b200.hparams.hp import Hparams
    from vae_gslm_b200.models.speech.lvtr import LVTR
    model = LVTR(Hparams.from_dict(copy.deepcopy(golden["config"])), input_dim=golden["n_mels"])
    model.load_state_dict(golden["state_dict"], strict=False)
    return model, ParamArena(model, bf16_shadow=False)


def test_arena_layout(golden):
    model, arena = _build(golden)
    params = list(model.parameters())
    # every parameter is a view into exactly one arena, 256-byte aligned, values preserved
    seen = 0
    for grp in arena.groups:
        base = grp.p.data_ptr()
        for p, off in zip(grp.params, grp.offsets):
            assert p.data_ptr() == base + 4 * off and off % 64 == 0
            assert p.grad is not None and p.grad.data_ptr() == grp.g.data_ptr() + 4 * off
            seen += 1
    assert seen == len(params)
    assert all(p.ndim != 1 for p in arena.decay.params) and all(p.ndim == 1 for p in arena.nodecay.params)
    sd = model.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in golden["state_dict"].items())
    # buckets are contiguous, disjoint and cover every parameter once
    buckets = arena.buckets(bucket_bytes=1 << 16)
    assert len(buckets) > 3
    assert sum(len(m) for _, m in buckets) == len(params)
    assert sum(b.numel() for b, _ in buckets) == arena.total_numel()
    # only non-direct gradients need zeroing; direct ones are the big transformer matrices, the conv stacks' 1x1
    # convolutions and the transformer layers' biases / RMSNorm scales (written by the GEMM / colsum / norm kernels)
    assert all(p.dim() in (1, 2, 3) for p in arena.direct) and len(arena.direct) >= 2 * 4 + 1
    assert sum(p.dim() == 1 for p in arena.direct) >= 2


def test_compact_checkpoint_roundtrip_and_resume(golden, tmp_path):
    """trainers/speech/lvtr.py save_checkpoint / load_checkpoint: the reference's compact format (state_dict under
    last-cpt.ckpt + hp.yaml, inference/inferer.py:17-28) plus the arena's AdamW state for resuming."""
    from vae_gslm_b200.hparams.hp import Hparams
    from vae_gslm_b200.trainers.speech.lvtr import load_checkpoint, save_checkpoint
    model, arena = _build(golden)
    hp = Hparams.from_dict({"model": copy.deepcopy(golden["config"])})
    torch.manual_seed(0)
    for grp in arena.groups:
        grp.m.normal_()
        grp.v.uniform_()
    arena.step_count = 17
    save_checkpoint(model, hp, str(tmp_path), arena=arena, global_step=1234)
    on_disk = torch.load(tmp_path / "last-cpt.ckpt", weights_only=True)
    # (the golden fixture holds the trainable tensors; the file also has the diffusion schedule buffers, as the reference's)
    assert set(on_disk) == set(model.state_dict()) and set(on_disk) >= set(golden["state_dict"])
    assert all(torch.equal(on_disk[k], v) for k, v in golden["state_dict"].items())
    assert Hparams.from_yamlfile(str(tmp_path / "hp.yaml")) == hp
    model2, arena2 = _build(golden)
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(1.0)
    assert load_checkpoint(model2, str(tmp_path), arena=arena2) == 1234 and arena2.step_count == 17
    for g1, g2 in zip(arena.groups, arena2.groups):
        assert torch.equal(g1.p, g2.p)
        for p, o in zip(g1.params, g1.offsets):          # (alignment gaps between slices are not state)
            sl = slice(o, o + p.numel())
            assert torch.equal(g1.m[sl], g2.m[sl]) and torch.equal(g1.v[sl], g2.v[sl])
    # parameters still alias the arena after loading
    p0 = next(model2.parameters())
    assert p0.data_ptr() in {g.p.data_ptr() + 4 * o for g in arena2.groups for o in g.offsets}


def test_arena_overwrite_calibration():
    """arena.py's first-writer-overwrite protocol, host side: a parameter written ONLY by direct sites (autograd sees
    None) joins the overwrite set, one that autograd accumulates into does not; the clear segments cover everything but
    the big overwrite slices; beta is 0 exactly once per step for those."""
    from vae_gslm_b200.arena import ParamArena

    class Direct(torch.autograd.Function):          # what ops._wgrad does: write the slice itself, hand autograd None
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            ctx.w = w
            return x @ w.t()

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            beta = ctx.w._vg_arena.wgrad_beta(ctx.w)
            ctx.w._vg_main_grad.mul_(beta).add_(g.t() @ x)
            return g @ w, None

    model = torch.nn.ModuleDict({"a": torch.nn.Linear(64, 256), "b": torch.nn.Linear(256, 16), "c": torch.nn.Linear(16, 8)})
    arena = ParamArena(model, bf16_shadow=False)
    arena._BIG = 1024
    wa, wb = model["a"].weight, model["b"].weight
    x = torch.randn(5, 64)

    def step():
        h = Direct.apply(x, wa)                                     # direct write
        h = torch.nn.functional.linear(h, wb, model["b"].bias)      # autograd accumulation
        model["c"](h).sum().backward()

    arena.calibrate_next()
    step()
    arena.finish_calibration()
    assert id(wa) in arena._overwrite and id(wb) not in arena._overwrite and not arena._cal_handles
    segs = arena._segments
    assert segs is not None
    cleared = torch.zeros(arena.decay.numel, dtype=torch.bool)
    off, ln, n = segs[0]
    for o, l in zip(off.tolist(), ln.tolist()):
        cleared[o:o + l] = True
    where = {id(p): k for k, p in enumerate(arena.decay.params)}
    oa, na = arena.decay.slice_of(where[id(wa)])
    ob, nb = arena.decay.slice_of(where[id(wb)])
    assert not cleared[oa:oa + na].any() and cleared[ob:ob + nb].all() and n == len(off)
    # per-step protocol (zero_grad itself needs the CUDA library: emulate it)
    expect = None
    for it in range(2):
        arena._written = set()
        for o, l in zip(off.tolist(), ln.tolist()):
            arena.decay.g[o:o + l] = 0
        arena.nodecay.g.zero_()
        step()
        ga = wa.grad.clone()
        if expect is None:
            expect = ga
        assert torch.allclose(ga, expect, rtol=1e-6, atol=1e-6)       # no stale gradient survives un-cleared slices
    assert arena.wgrad_beta(wa) == 1.0 and arena.wgrad_beta(wb) == 1.0   # already written this step / not in the set
    arena._written = set()
    assert arena.wgrad_beta(wa) == 0.0 and arena.wgrad_beta(wa) == 1.0


def _worker(rank, world, port, fixture_path, result_q, compress=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vae_gslm_b200.dp import GradReducer
        golden = torch.load(fixture_path, map_location="cpu", weights_only=False)
        model, arena = _build(golden)
        reducer = GradReducer(arena, bucket_bytes=1 << 16, compress_bf16=compress)
        params = list(model.parameters())

        def backward(scale):
            arena.zero_grad()
            for p in arena.direct:                           # on the GPU path the wgrad GEMM overwrites these (beta=0);
                p.grad.zero_()                               # here autograd accumulates into them, so clear them too
            loss = sum((p * scale).sum() for p in params)
            loss.backward()                                  # post-accumulate hooks fire for the hooked parameters
            for p in arena.direct:                           # what ops._wgrad does after writing the arena slice
                arena.grad_ready(p)

        # micro-batch that does NOT close the window: no communication, local gradients stay
        reducer.prepare(last_micro_batch=False)
        backward(float(rank + 1))
        reducer.finish()
        local_only = all(torch.allclose(p.grad, torch.full_like(p.grad, float(rank + 1))) for p in params)
        # closing micro-batch: every gradient becomes the mean over ranks
        reducer.prepare(last_micro_batch=True)
        backward(float(rank + 1))
        launched_during_backward = sum(reducer._launched)
        reducer.finish()
        mean = sum(range(1, world + 1)) / world
        averaged = all(torch.allclose(p.grad, torch.full_like(p.grad, mean)) for p in params)
        result_q.put((rank, local_only, averaged, launched_during_backward, len(reducer.buckets)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("compress", [False, True])
def test_grad_reducer_world2_gloo(golden, tmp_path, compress):
    fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lvtr_small.pt")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fixture, q, compress)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, local_only, averaged, launched, nbuckets in results:
        assert local_only, f"rank {rank}: gradients were reduced on a non-boundary micro-batch"
        assert averaged, f"rank {rank}: gradients are not the mean over ranks"
        assert launched == nbuckets, "every bucket should be launched from the backward hooks (overlap), none in finish()"
