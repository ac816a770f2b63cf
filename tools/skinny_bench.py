"""vg_skinny_linear against the general tcgen05 GEMM (skinny-M plan) and cuBLAS on the shapes of one generation step,
graph-timed (16 layers' worth of launches per replay, distinct weights per launch so that W streams from HBM).
usage: python tools/skinny_bench.py [B ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib, ops

dev = torch.device("cuda", 0)
_lib.load()
bf = torch.bfloat16
Bs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [16, 64, 128, 256]
PLAIN = "--plain" in sys.argv
L = 16
for B in Bs:
    for (N, K, tag) in [(3072, 1024, "qkv"), (1024, 1024, "out"), (4096, 1024, "ffn1"), (1024, 4096, "ffn2")]:
        x = torch.randn(B, K, device=dev).to(bf)
        ws = [(torch.randn(N, K, device=dev) / K ** 0.5).to(bf) for _ in range(L)]
        res = torch.randn(B, N, device=dev).to(bf)

        def timed(fn):
            for _ in range(2):
                for w in ws:
                    fn(w)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for w in ws:
                    fn(w)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / (5 * L)
        sk = timed(lambda w: ops.skinny_linear(x, w, None, ops.ACT_NONE, None if PLAIN else res))
        gm = timed(lambda w: ops.gemm(x, w, residual=res))
        cb = timed(lambda w: torch.addmm(res, x, w.t()))
        floor = N * K * 2 / 6547.5e3
        print(f"B={B:4d} {tag:5s} [{N}x{K}] + residual: skinny {sk:6.2f} us   gemm_tc {gm:6.2f} us   cuBLAS {cb:6.2f} us   weight-stream floor {floor:5.2f} us", flush=True)
