"""Data-parallel correctness on hardware (reference DDP mean semantics scripts/train.py:93-95):
  torchrun --nproc-per-node 2 tools/dp_check.py
(1) the gradient every rank holds after the bucketed, overlapped NCCL all-reduce of its LOCAL micro-batch equals the
    gradient of ONE process on the concatenated batch divided by the world size;
(2) after k optimizer steps on different local batches all ranks hold bit-identical parameters."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from vae_gslm_b200 import _lib
from vae_gslm_b200.arena import ParamArena
from vae_gslm_b200.dp import GradReducer
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR
from vae_gslm_b200.trainers.speech.lvtr import TrainStep
from vae_gslm_b200.training_lib.trainer import init_weights

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
_lib.load()
B, T = 2, 256


def build():
    torch.manual_seed(0)
    m = LVTR(Hparams.from_yamlfile(bench.CFG).model, input_dim=80)
    m.apply(init_weights)
    return m.to(dev).set_compute_dtype(torch.bfloat16)


def draws(r):
    return {k: v.to(dev) for k, v in bench._rng(B, T, seed=4321 + r).items()}


def inject(model, d):
    inner = model.forward
    model.forward = lambda x, _f=inner, **kw: _f(x, **kw, **d)


def flat(arena):
    return torch.cat([g.g.detach().float() for g in arena.groups])


# ---- (1) all-reduced local gradients vs the single-process gradient of the concatenated batch
model = build()
inject(model, draws(rank))
arena = ParamArena(model, weight_decay=0.1)
reducer = GradReducer(arena, bucket_bytes=64 << 20)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(B, T, rank).items()}
step = TrainStep(model, arena, reducer, batch, lr=0.0, kld_weight=bench.KW, use_cuda_graph=False)
step(lr=0.0)
torch.cuda.synchronize()
g_dp = flat(arena)
gathered = [torch.zeros_like(g_dp) for _ in range(world)]
dist.all_gather(gathered, g_dp)
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    m1 = build()
    cat_draws = {k: torch.cat([draws(r)[k] for r in range(world)], 0) for k in draws(0)}
    inject(m1, cat_draws)
    a1 = ParamArena(m1, weight_decay=0.1)
    big = {k: torch.cat([bench.synthetic_batch(B, T, r)[k].to(dev) for r in range(world)], 0) for k in batch}
    r1 = GradReducer(a1, process_group=None)
    r1.world = 1                               # a private, NON-communicating reducer for the single-process side
    s1 = TrainStep(m1, a1, r1, big, lr=0.0, kld_weight=bench.KW, use_cuda_graph=False)
    s1(lr=0.0)
    torch.cuda.synchronize()
    g_one = flat(a1) / world
    err = float((g_dp - g_one).norm() / g_one.norm())
    print(f"[dp_check] all-reduced gradient identical on all {world} ranks: {same}; vs single-process gradient of the "
          f"concatenated batch / {world}: Frobenius-relative {err:.3e} (bf16 step, bound 2e-2)", flush=True)
    assert same and err < 2e-2
dist.barrier()

# ---- (2) k optimizer steps on different local batches (CUDA-graph step): parameters stay bit-identical across ranks
step2 = TrainStep(model, arena, reducer, batch, lr=2e-4, kld_weight=bench.KW, use_cuda_graph=True)
assert step2.graph is not None, step2.capture_error
for i in range(6):
    step2.load({k: v.to(dev) for k, v in bench.synthetic_batch(B, T, 100 * i + rank).items()})
    step2(lr=2e-4)
torch.cuda.synchronize()
p = torch.cat([g.p.detach() for g in arena.groups])
ps = [torch.zeros_like(p) for _ in range(world)]
dist.all_gather(ps, p)
if rank == 0:
    ident = all(torch.equal(ps[0], q) for q in ps)
    print(f"[dp_check] parameters bit-identical across {world} ranks after 6 graph-replayed optimizer steps on different "
          f"local batches: {ident} (step count {arena.step_count})", flush=True)
    assert ident
step2.graph = None
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
print(f"[dp_check] rank {rank}: process group destroyed cleanly", flush=True)
