"""Kernel-only timing of vg_attn_fwd / vg_attn_bwd through the C ABI (no autograd, no gradient accumulation): CUDA events
around 20 back-to-back launches.  The backward call is everything the ABI runs: delta kernel, dQ-accumulator memset, the
main kernel and the fp32→bf16 dQ conversion."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops

shapes = [(8, 1000, 16), (8, 1000, 16), (8, 640, 16), (2, 3000, 16)]
if len(sys.argv) > 3:
    shapes = [tuple(int(v) for v in sys.argv[1:4])]
for (B, T, H) in shapes:
    D, HD, C3 = 64, H * 64, 3 * H * 64
    qkv = (0.5 * torch.randn(B, T, C3, device="cuda")).to(torch.bfloat16)
    q, k, v = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
    out = torch.empty(B, T, HD, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda", dtype=torch.float32)
    dout = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = dqkv[..., :HD], dqkv[..., HD:2 * HD], dqkv[..., 2 * HD:]
    slopes = torch.tensor(ops.alibi_slopes(H), device="cuda")
    kv_len = torch.full((B,), T, device="cuda", dtype=torch.int32)
    ws = L.workspace(L.load().vg_attn_bwd_workspace(B, H, T, T, D), qkv.device)
    scale = 1.0 / math.sqrt(D)

    def fwd():
        L.call("vg_attn_fwd", L.ptr(q), L.ptr(k), L.ptr(v), C3, C3, L.ptr(out), HD, L.ptr(lse), L.ptr(kv_len),
               L.ptr(slopes), B, H, T, T, D, 0, 0, 0, scale, L.dtype_id(qkv.dtype), L.stream())

    def bwd():
        L.call("vg_attn_bwd", L.ptr(dout), HD, L.ptr(q), L.ptr(k), L.ptr(v), C3, C3, L.ptr(out), HD, L.ptr(lse),
               L.ptr(dq), L.ptr(dk), L.ptr(dv), C3, C3, L.ptr(kv_len), L.ptr(slopes), B, H, T, T, D, 0, scale,
               L.dtype_id(qkv.dtype), L.ptr(ws), ws.numel(), L.stream())

    res = []
    for fn in (fwd, bwd):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1e3)
    fl = 4.0 * B * H * T * T * 64 / 2
    print(f"raw B={B} T={T} H={H}  fwd {res[0]:7.1f} us ({fl / res[0] / 1e6:6.1f} TFLOP/s)   bwd {res[1]:7.1f} us "
          f"({2.5 * fl / res[1] / 1e6:6.1f} TFLOP/s)", flush=True)
