"""Per-shape timing of the tcgen05 GEMM on the shapes of one VAE-GSLM training step (M = B*T rows), against
cuBLAS (torch.matmul) on the same GPU.  Usage: python tools/gemm_bench.py [M]   (default 8000 = 8 x 1000 frames)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops

dev = "cuda"
args = [a for a in sys.argv[1:] if not a.startswith("--")]
M = int(args[0]) if args else 8000
WGRAD_ONLY = "--wgrad-only" in sys.argv
QUICK = "--quick" in sys.argv or WGRAD_ONLY
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
    return e0.elapsed_time(e1) / n


def report(name, flops, ours, ref):
    print(f"{name:34s} ours {ours*1e3:8.1f} us {flops/ours/1e9:7.1f} TF | cublas {ref*1e3:8.1f} us {flops/ref/1e9:7.1f} TF"
          f" | ratio {ref/ours:5.2f}", flush=True)


x1k = torch.randn(M, 1024, device=dev).to(bf)
x4k = torch.randn(M, 4096, device=dev).to(bf)
mask = torch.ones(M, dtype=torch.uint8, device=dev)
SHAPES = [(3072, 1024, "qkv"), (1024, 1024, "out_proj"), (4096, 1024, "ffn1"), (1024, 4096, "ffn2")]
if not QUICK:
    SHAPES += [(2048, 1024, "spliters"), (520, 1024, "head"), (200, 1024, "logits"), (1024, 64, "stack_in"),
               (2048, 544, "conv_up"), (512, 2048, "conv_down")]
for (N, K, tag) in SHAPES:
    x = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf)
    dy = torch.randn(M, N, device=dev).to(bf)
    fl = 2.0 * M * N * K
    acc = torch.zeros(N, K, device=dev)
    report(f"wgrad+= {tag} [{N},{M}]x[{M},{K}] f32 beta=1", fl,
           timeit(lambda: ops.gemm(dy, x, trans_a=True, trans_b=False, out=acc, beta=1.0)),
           timeit(lambda: dy.t() @ x))
    if WGRAD_ONLY:
        continue
    report(f"fwd   {tag} [{M},{K}]x[{N},{K}]T", fl, timeit(lambda: ops.gemm(x, w)), timeit(lambda: x @ w.t()))
    report(f"dgrad {tag} [{M},{N}]x[{N},{K}]", fl, timeit(lambda: ops.gemm(dy, w, trans_b=False)),
           timeit(lambda: dy @ w))
    report(f"wgrad {tag} [{N},{M}]x[{M},{K}]", fl,
           timeit(lambda: ops.gemm(dy, x, trans_a=True, trans_b=False, out_dtype=torch.float32)),
           timeit(lambda: dy.t() @ x))
if WGRAD_ONLY:
    sys.exit(0)
# fused-epilogue variants actually used by the layer
w1 = (torch.randn(4096, 1024, device=dev) / 32).to(bf)
b1 = torch.randn(4096, device=dev)
pre = torch.empty(M, 4096, device=dev, dtype=bf)
report("fwd ffn1 +bias+gelu+preact", 2.0 * M * 4096 * 1024,
       timeit(lambda: ops.gemm(x1k, w1, bias=b1, act=ops.ACT_GELU, preact=pre)), timeit(lambda: x1k @ w1.t()))
w2 = (torch.randn(1024, 4096, device=dev) / 64).to(bf)
b2 = torch.randn(1024, device=dev)
report("fwd ffn2 +bias+res+mask", 2.0 * M * 4096 * 1024,
       timeit(lambda: ops.gemm(x4k, w2, bias=b2, residual=x1k, row_mask=mask)), timeit(lambda: x4k @ w2.t()))
dy1k = torch.randn(M, 1024, device=dev).to(bf)
report("dgrad ffn2 * gelu'(pre)", 2.0 * M * 4096 * 1024,
       timeit(lambda: ops.gemm(dy1k, w2, trans_b=False, dact_src=pre, dact=ops.ACT_GELU)), timeit(lambda: dy1k @ w2))
report("fwd ffn1 +bias+gelu+gelu'(pre) saved", 2.0 * M * 4096 * 1024,
       timeit(lambda: ops.gemm(x1k, w1, bias=b1, act=ops.ACT_GELU, preact=pre, preact_is_grad=True)),
       timeit(lambda: x1k @ w1.t()))
report("dgrad ffn2 * saved derivative", 2.0 * M * 4096 * 1024,
       timeit(lambda: ops.gemm(dy1k, w2, trans_b=False, dact_src=pre, dact=ops.ACT_MULT)), timeit(lambda: dy1k @ w2))
