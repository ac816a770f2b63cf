"""a few vg_decode_linear launches of one shape — target of `ncu --set full` captures."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops
B, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 4096, 1024)
bf = torch.bfloat16
x = torch.randn(B, K, device="cuda").to(bf)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(bf)
ws = ops.decode_linear_workspace(B, N, "cuda")
out = torch.empty(B, N, device="cuda", dtype=bf)
for _ in range(6):
    ops.decode_linear(x, w, ws, out=out, overlap=False)
torch.cuda.synchronize()
