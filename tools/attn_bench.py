"""Micro-benchmark of the attention kernels (run on the GPU box): forward / backward time and useful TFLOP/s
(causal half only: 4·B·H·T²·D/2 forward, 2.5x that backward) for the SIMT and tcgen05 backends."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops

dev = "cuda"
for (B, T, H) in [(8, 1000, 16), (8, 1000, 16), (8, 640, 16), (2, 3000, 16)]:
    qkv = (0.5 * torch.randn(B, T, 3 * H * 64, device=dev)).to(torch.bfloat16).requires_grad_(True)
    slopes = torch.tensor(ops.alibi_slopes(H), device=dev)
    lengths = torch.full((B,), T, device=dev, dtype=torch.int32)
    flops_f = 4.0 * B * H * T * T * 64 / 2
    for backend in ("tcgen05", "simt"):
        L.set_attention_backend(backend)
        o = ops.attention(qkv, H, lengths, slopes)
        go = torch.randn_like(o)
        for _ in range(2):
            o = ops.attention(qkv, H, lengths, slopes)
            o.backward(go)
        torch.cuda.synchronize()
        n = 10 if backend == "tcgen05" else 2
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        outs = [ops.attention(qkv, H, lengths, slopes) for _ in range(n)]
        e[1].record()
        for o in outs:
            o.backward(go)
        e[2].record()
        torch.cuda.synchronize()
        tf, tb = e[0].elapsed_time(e[1]) / n, e[1].elapsed_time(e[2]) / n
        print(f"B={B} T={T} H={H} {backend:8s} fwd {tf*1e3:8.1f} us ({flops_f/tf/1e9:7.1f} TFLOP/s)  "
              f"bwd {tb*1e3:8.1f} us ({2.5*flops_f/tb/1e9:7.1f} TFLOP/s)", flush=True)
L.set_attention_backend("auto")
