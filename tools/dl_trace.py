"""phase timeline (clock64 deltas of CTA 0) of vg_decode_linear for a few shapes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops
bf = torch.bfloat16
trace = torch.zeros(48, dtype=torch.int64, device="cuda")
L.call("vg_debug_decode_linear_trace", L.ptr(trace))
names = ["W issue", "pdl wait", "x issue", "cp.async wait+sync", "mma+sync", "epilogue"]
for (B, N, K) in ((1, 4096, 1024), (1, 1024, 4096), (64, 4096, 1024), (64, 1024, 4096), (64, 3072, 1024), (256, 4096, 1024)):
    x = torch.randn(B, K, device="cuda").to(bf)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(bf)
    ws = ops.decode_linear_workspace(B, N, "cuda")
    out = torch.empty(B, N, device="cuda", dtype=bf)
    for _ in range(3):
        ops.decode_linear(x, w, ws, out=out, overlap=False)
    torch.cuda.synchronize()
    t = trace.cpu().tolist()
    d = [t[i + 1] - t[i] for i in range(6)]
    print("   per-warp (cycles after kernel start): W landed", [v - t[0] for v in t[16:24]], "x landed", [v - t[0] for v in t[24:32]],
          "mma done", [v - t[0] for v in t[8:16]], flush=True)
    print("   per-element stamps after reads:", [t[40 + j] - t[32] for j in range(8)], flush=True)
    print("   epilogue detail: sync->loop", t[34] - t[5], "reads", t[32] - t[34], "math", t[33] - t[32], "store+rest", t[6] - t[33], flush=True)
    print(f"B={B} N={N} K={K}: total {t[6] - t[0]} cycles | " + " | ".join(f"{n} {v}" for n, v in zip(names, d)), flush=True)
