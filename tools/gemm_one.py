"""one tcgen05 GEMM shape, a few launches — the target of `ncu --set full` captures (profiles/).
usage: python tools/gemm_one.py M N K [plain|gelu|mult|wgrad]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8000, 4096, 1024)
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
bf = torch.bfloat16
x = torch.randn(M, K, device="cuda").to(bf)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(bf)
b = torch.randn(N, device="cuda")
pre = torch.empty(M, N, device="cuda", dtype=bf)
dy = torch.randn(M, N, device="cuda").to(bf)
acc = torch.zeros(N, K, device="cuda")
for _ in range(6):
    if mode == "plain":
        ops.gemm(x, w)
    elif mode == "gelu":
        ops.gemm(x, w, bias=b, act=ops.ACT_GELU, preact=pre, preact_is_grad=True)
    elif mode == "mult":       # dgrad of the layer after an activation: [M,K_out] x [K_out,N] * saved derivative
        ops.gemm(x, w.t().contiguous(), trans_b=False, dact_src=pre, dact=ops.ACT_MULT)
    elif mode == "wgrad":
        ops.gemm(dy, x, trans_a=True, trans_b=False, out=acc, beta=1.0)
torch.cuda.synchronize()
