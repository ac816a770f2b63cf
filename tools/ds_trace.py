"""Per-phase clock trace of the persistent decode-step kernel, every CTA: per phase type, the time a CTA spends from
leaving barrier p-1 to entering barrier p ("work": mean / max over CTAs — the max is the phase's critical path), the time
inside the barrier ("wait": the min over CTAs is the barrier's own latency), and for GEMM phases the stages of the slowest CTA.
usage: python tools/ds_trace.py [B] [barrier mode] [prompt length]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vae_gslm_b200 import _lib
from vae_gslm_b200.decode_step import DecodeStepEngine
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR
from vae_gslm_b200.training_lib.trainer import init_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
P = int(sys.argv[3]) if len(sys.argv) > 3 else 400
dev = torch.device("cuda", 0)
_lib.load()
torch.manual_seed(0)
model = LVTR(Hparams.from_yamlfile(bench.CFG).model, input_dim=80)
model.apply(init_weights)
model = model.to(dev).set_compute_dtype(torch.bfloat16).eval()
model.use_decode_engine = False
model.transformer[0].cache_len_hint = P + 64
prompt = torch.cat([torch.randint(0, 200, (B, P, 1), device=dev).float(), torch.randn(B, P, 4, device=dev)], -1)
kv = model.step(prompt, past_kv=None, temperature=0.0, push_init_state=True, greedy=True)["kv"]
eng = DecodeStepEngine(model, B, dev, barrier_mode=mode)
u = torch.randn(B, 64, device=dev)
for _ in range(3):
    eng.run(u, kv)
G, NP = eng.grid, eng.NP
eng.trace = torch.zeros(G * NP * 8, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
eng.run(u, kv)
e1.record()
torch.cuda.synchronize()
print(f"B={B} mode={mode} Tk={kv[0].cache.length}: one launch (with u copy / pos fill) {e0.elapsed_time(e1) * 1e3:.1f} us")
tr = eng.trace.view(G, NP, 8).cpu().double()
work = tr[:, 1:NP - 1, 0] - tr[:, 0:NP - 2, 1]          # [G, phases 1..NP-2]
wait = tr[:, 1:NP - 1, 1] - tr[:, 1:NP - 1, 0]
names = [n.rstrip("0123456789").rstrip("_") for n in eng.phase_names[1:NP - 1]]
print(f"total (CTA 0, first to last barrier exit): {int(tr[0, NP - 2, 1] - tr[0, 0, 1])} cycles")
print(f"{'phase':6s} {'n':>3s} {'work mean':>10s} {'work max':>10s} {'wait min':>10s} {'phase len':>10s} | slowest CTA: X published, all mma issued, acc ready, first tmem read, first chunk reduced, epilogue done (cycles after barrier exit)")
for k in dict.fromkeys(names):
    idx = [i for i, n in enumerate(names) if n == k]
    w, wt = work[:, idx], wait[:, idx]
    plen = (w + wt).mean()
    slow = w.argmax(0)                                  # slowest CTA per phase instance
    st = []
    for slot in (7, 3, 4, 6, 2, 5):
        vals = [float(tr[int(slow[j]), idx[j] + 1, slot] - tr[int(slow[j]), idx[j], 1]) for j in range(len(idx))]
        st.append(sum(vals) / len(vals))
    print(f"{k:6s} {len(idx):3d} {float(w.mean()):10.0f} {float(w.max(0).values.mean()):10.0f} {float(wt.min(0).values.mean()):10.0f} "
          f"{float(plen):10.0f} | " + "  ".join(f"{v:8.0f}" for v in st))
