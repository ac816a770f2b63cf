"""clock64 phase trace of one CTA of the tcgen05 attention kernels (linear block index from argv[4], default 0: forward =
the heaviest query tile; backward = key tile 0, which loops over every query tile).  Prints per-iteration phase stamps."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops
B, T, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 1000, 16)
blocks = [int(v) for v in sys.argv[4:]] or [0]
qkv = (0.5 * torch.randn(B, T, 3 * H * 64, device="cuda")).to(torch.bfloat16).requires_grad_(True)
slopes = torch.tensor(ops.alibi_slopes(H), device="cuda")
lengths = torch.full((B,), T, device="cuda", dtype=torch.int32)
for _ in range(2):
    o = ops.attention(qkv, H, lengths, slopes)
    o.backward(torch.randn_like(o))
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
g = torch.randn_like(o)


def show(name, t, labels):
    marks = [int(v) for v in t[30]]
    t = t[:30]
    n = int((t[:, 0] != 0).sum())
    t0 = marks[0]
    print(f"{name}: {n} iterations; cycles since CTA entry: " + " | ".join(labels))
    for i in range(n):
        print(f"  it {i:2d}: " + " ".join(f"{int(v) - t0:7d}" if int(v) else "      -" for v in t[i]))
    print("  marks (entry, setup done, loop done, epilogue done, exit, [bwd: reductions complete]): "
          + " ".join(str(m - t0) if m else "-" for m in marks[:6]))
    if n > 1:
        print(f"  cycles per iteration: {(int(t[n - 1, 0]) - int(t[0, 0])) / (n - 1):.0f}")


for blk in blocks:
    trace.zero_()
    trace[255] = blk
    L.call("vg_debug_attn_trace", L.ptr(trace))
    o = ops.attention(qkv, H, lengths, slopes)
    torch.cuda.synchronize()
    fwd = trace.cpu().view(-1, 8).clone()
    trace.zero_()
    trace[255] = blk
    o.backward(g)
    torch.cuda.synchronize()
    bwd = trace.cpu().view(-1, 8).clone()
    L.call("vg_debug_attn_trace", None)
    print(f"==== block {blk}")
    show("forward", fwd, ["softmax: S ready", "pass-1 done", "P published", "O_j ready", "O folded",
                          "mma: S(j) issued", "P(j-1)V issued", "P ready"])
    show("backward", bwd, ["elementwise: S,dP ready", "math done", "P,dS published", "dQ(it-1) drained",
                           "mma: dQ issued", "S,dP(it) issued", "P,dS ready", "dV,dK issued"])
