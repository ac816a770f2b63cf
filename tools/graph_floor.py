"""micro-benchmark: per-kernel cost of launch chains replayed from a CUDA graph (decode-path design input)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
L.load()
bf = torch.bfloat16


def bench(name, body, n_kernels, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:60s} {us:8.1f} us / replay   {us / n_kernels:6.2f} us per kernel", flush=True)


cnt = torch.zeros(1, dtype=torch.int32, device=dev)
bench("100 x add_i32 (1 thread)", lambda: [L.call("vg_add_i32", L.ptr(cnt), 1, L.stream()) for _ in range(100)], 100)
for B in (1, 64):
    for (N, K) in ((4096, 1024), (1024, 4096), (1024, 1024)):
        x = torch.randn(B, K, device=dev).to(bf)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf)
        ws = ops.decode_linear_workspace(B, N, dev)
        out = torch.empty(B, N, device=dev, dtype=bf)
        for ov in (False, True):
            bench(f"100 x decode_linear B={B} N={N} K={K} overlap={ov}",
                  lambda: [ops.decode_linear(x, w, ws, out=out, overlap=ov) for _ in range(100)], 100)
x = torch.randn(1, 1024, device=dev).to(bf)
w = (torch.randn(4096, 1024, device=dev) / 32).to(bf)
ws = ops.decode_linear_workspace(1, 4096, dev)
out = torch.empty(1, 4096, device=dev, dtype=bf)


def alt():
    for _ in range(50):
        ops.decode_linear(x, w, ws, out=out, overlap=True)
        L.call("vg_add_i32", L.ptr(cnt), 1, L.stream())


bench("50 x (decode_linear B=1 4096x1024 + add_i32)", alt, 100)
