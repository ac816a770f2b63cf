"""tcgen05 GEMM on the shapes of one cached generation step (M = decode batch) against cuBLAS and the weight-streaming
floor 2·N·K bytes / HBM bandwidth.  usage: python tools/gemm_skinny_bench.py [M ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops

dev, bf = "cuda", torch.bfloat16
Ms = [int(a) for a in sys.argv[1:]] or [64, 128, 256]


def timeit(fn, n=10):
    """16 calls (one per weight matrix) captured into a CUDA graph: device time per call, no host launch overhead"""
    for _ in range(16):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(16):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (16 * n) * 1e3


# 16 different weight matrices per shape, cycled, so that the weights come from HBM as in a real step (16 layers)
for M in Ms:
    for (N, K, mode) in ((3072, 1024, "bf16"), (1024, 1024, "f32+="), (4096, 1024, "gelu"), (1024, 4096, "f32+=")):
        Ws = [(torch.randn(N, K, device=dev) / K ** 0.5).to(bf) for _ in range(16)]
        x = torch.randn(M, K, device=dev).to(bf)
        bias = torch.randn(N, device=dev)
        acc = torch.zeros(M, N, device=dev)
        i = [0]

        def ours():
            w = Ws[i[0] % 16]
            i[0] += 1
            if mode == "bf16":
                ops.gemm(x, w)
            elif mode == "gelu":
                ops.gemm(x, w, bias=bias, act=ops.ACT_GELU)
            else:
                ops.gemm(x, w, out=acc, beta=1.0)

        def cublas():
            w = Ws[i[0] % 16]
            i[0] += 1
            torch.matmul(x, w.t())

        t1, t2 = timeit(ours), timeit(cublas)
        print(f"M={M:4d} [{N}x{K}] {mode:6s}: ours {t1:6.1f} us  cuBLAS {t2:6.1f} us  weight-stream floor {2 * N * K / 6547.5e3:5.2f} us", flush=True)
