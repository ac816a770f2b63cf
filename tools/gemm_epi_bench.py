"""tcgen05 GEMM with the fused epilogues of the training step (M = 8000): plain vs residual / row mask / bias / stored
derivative, per shape, CUDA events over 20 launches.  usage: python tools/gemm_epi_bench.py [M]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops
from vae_gslm_b200._lib import ACT_GELU, ACT_MULT, ACT_NONE

dev = "cuda"
_a = [a for a in sys.argv[1:] if a.isdigit()]
M = int(_a[0]) if _a else 8000
bf = torch.bfloat16


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


COLD = "--cold" in sys.argv
if COLD:
    # every launch sees a cold L2, as inside the training step (the working set of a layer does not stay resident):
    # a 512 MB memset between launches, one event pair per launch, a spin kernel in front so the host runs ahead
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, n=10):                                    # noqa: F811
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda._sleep(int(0.01 * 1.9e9))
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize()
        return sum(e0.elapsed_time(e1) for e0, e1 in ev) / n * 1e3

    def cublas(a, b):
        return a @ b

mask = (torch.rand(M, device=dev) > 0.1).to(torch.uint8)
for (N, K, tag) in [(1024, 1024, "out_proj"), (1024, 4096, "ffn2"), (4096, 1024, "ffn1"), (3072, 1024, "qkv")]:
    x = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf)
    res = torch.randn(M, N, device=dev).to(bf)
    bias = torch.randn(N, device=dev)
    dsrc = torch.rand(M, K, device=dev).to(bf)
    dy = torch.randn(M, N, device=dev).to(bf)
    fl = 2.0 * M * N * K
    rows = [
        ("cuBLAS fwd (torch.matmul)", lambda: x @ w.t()),
        ("fwd plain", lambda: ops.gemm(x, w)),
        ("fwd + residual", lambda: ops.gemm(x, w, residual=res)),
        ("fwd + residual + mask", lambda: ops.gemm(x, w, residual=res, row_mask=mask, mask_first=True)),
        ("fwd + bias + residual + mask", lambda: ops.gemm(x, w, bias=bias, residual=res, row_mask=mask)),
        ("dgrad plain", lambda: ops.gemm(dy, w, trans_b=False)),
        ("dgrad x stored derivative", lambda: ops.gemm(dy, w, trans_b=False, dact_src=dsrc, dact=ACT_MULT)),
    ]
    if tag == "ffn1":
        pre = torch.empty(M, N, device=dev, dtype=bf)
        rows.append(("fwd + bias + GELU + derivative", lambda: ops.gemm(x, w, bias=bias, act=ACT_GELU, preact=pre, preact_is_grad=True)))
    for name, fn in rows:
        us = timeit(fn)
        print(f"{tag:9s} {name:32s} {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
