"""a few attention fwd+bwd launches at one shape — target of `ncu --set full` captures."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops
B, T, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 1000, 16)
qkv = (0.5 * torch.randn(B, T, 3 * H * 64, device="cuda")).to(torch.bfloat16).requires_grad_(True)
slopes = torch.tensor(ops.alibi_slopes(H), device="cuda")
lengths = torch.full((B,), T, device="cuda", dtype=torch.int32)
for _ in range(3):
    o = ops.attention(qkv, H, lengths, slopes)
    o.backward(torch.randn_like(o))
torch.cuda.synchronize()
