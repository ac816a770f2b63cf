"""clock64 phase trace of CTA (0,0,0) of the tcgen05 attention kernels (forward: the heaviest query tile; backward: key
tile 0, which loops over every query tile).  Prints per-iteration phase durations in cycles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops
B, T, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 1000, 16)
qkv = (0.5 * torch.randn(B, T, 3 * H * 64, device="cuda")).to(torch.bfloat16).requires_grad_(True)
slopes = torch.tensor(ops.alibi_slopes(H), device="cuda")
lengths = torch.full((B,), T, device="cuda", dtype=torch.int32)
for _ in range(2):
    o = ops.attention(qkv, H, lengths, slopes)
    o.backward(torch.randn_like(o))
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
g = torch.randn_like(o)
L.call("vg_debug_attn_trace", L.ptr(trace))
o = ops.attention(qkv, H, lengths, slopes)
torch.cuda.synchronize()
fwd = trace.cpu().view(-1, 8).clone()
trace.zero_()
o.backward(g)
torch.cuda.synchronize()
bwd = trace.cpu().view(-1, 8).clone()
L.call("vg_debug_attn_trace", None)


def show(name, t, labels):
    n = int((t[:, 0] != 0).sum())
    t0 = int(t[:n][t[:n] != 0].min())
    print(f"{name}: {n} iterations; columns = cycles since the first stamp: " + " | ".join(labels))
    for i in range(n):
        print(f"  it {i:2d}: " + " ".join(f"{int(v) - t0:7d}" if int(v) else "      -" for v in t[i]))
    print(f"  cycles per iteration: {(int(t[n - 1].max()) - t0) / n:.0f}")


show("forward", fwd, ["softmax: S ready", "pass-1 done", "P published", "O_j ready", "O folded",
                      "mma: K landed", "S issued", "P+V ready"])
show("backward", bwd, ["elementwise: S,dP ready", "P,dS published", "dQ ready", "dQ drained",
                       "mma: Q,dO landed", "S,dP issued", "P,dS ready", "dV,dK,dQ issued"])
