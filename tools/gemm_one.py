"""one tcgen05 GEMM shape, a few launches — the target of `ncu --set full` captures (profiles/)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8000, 4096, 1024)
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
for _ in range(6):
    ops.gemm(x, w)
torch.cuda.synchronize()
