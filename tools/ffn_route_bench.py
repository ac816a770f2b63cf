"""The FFN of a generation step at 65..256 sequences: W2 with the fused bias / residual / mask epilogue (the product
route: 8..32 output tiles, each a serial pass over 64 k-blocks) against an experiment that splits the k-range across the
SMs — f32 accumulator initialised to residual + b2, vg_gemm's pure-accumulation split-K plan, mask on the way back to
bf16.  Graph-timed over 12 distinct layers' weights (they stream from HBM as in a generation step).  Result in
profiles/r02_decode.md section 3c: no gain once the step runs as a programmatic-dependent-launch chain.
usage: python tools/ffn_route_bench.py [rows ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib, ops

_lib.load()
dev = torch.device("cuda", 0)
bf = torch.bfloat16
D, F, LAYERS = 1024, 4096, 12
rows = [int(a) for a in sys.argv[1:]] or [128, 256]
w1 = [(torch.randn(F, D, device=dev) / 32).to(bf) for _ in range(LAYERS)]
w2 = [(torch.randn(D, F, device=dev) / 64).to(bf) for _ in range(LAYERS)]
b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / LAYERS


for M in rows:
    x = torch.randn(M, D, device=dev).to(bf)
    res = torch.randn(M, D, device=dev).to(bf)
    mask = (torch.rand(M, device=dev) > 0.2)
    h = torch.randn(M, F, device=dev).to(bf)
    acc = torch.empty(M, D, device=dev)
    y = torch.empty(M, D, device=dev, dtype=bf)
    m8 = mask.view(torch.uint8)

    def fused():
        with torch.no_grad():
            for l in range(LAYERS):
                ops.ffn(x, w1[l], b1, w2[l], b2, residual=res, row_mask=mask)

    def split_k():
        for l in range(LAYERS):
            hh = ops.gemm(x, w1[l], bias=b1, act=ops.ACT_GELU)
            a = torch.empty(M, D, device=dev)
            torch.add(res, b2, out=a)
            ops.gemm(hh, w2[l], out=a, beta=1.0)
            torch.mul(a, m8.view(M, 1), out=torch.empty(M, D, device=dev, dtype=bf))

    out = {}
    for flag, fn in ((False, fused), (True, split_k)):
        for pdl in (0, 3):
            with ops.pdl_mode(pdl):
                out[(flag, pdl)] = timed(fn)
    pieces = {
        "ffn1 +preact": lambda: [ops.gemm(x, w1[l], bias=b1, act=ops.ACT_GELU, preact=torch.empty(M, F, device=dev, dtype=bf), preact_is_grad=True) for l in range(LAYERS)],
        "ffn1": lambda: [ops.gemm(x, w1[l], bias=b1, act=ops.ACT_GELU) for l in range(LAYERS)],
        "ffn2 fused": lambda: [ops.gemm(h, w2[l], bias=b2, residual=res, row_mask=m8) for l in range(LAYERS)],
        "ffn2 split-K into f32": lambda: [ops.gemm(h, w2[l], out=acc, beta=1.0) for l in range(LAYERS)],
        "add(res, b2) -> f32": lambda: [torch.add(res, b2, out=acc) for l in range(LAYERS)],
        "mul(acc, mask) -> bf16": lambda: [torch.mul(acc, m8.view(M, 1), out=y) for l in range(LAYERS)],
    }
    print(f"rows {M}: us per layer  fused {out[(False, 0)]:.1f} (pdl {out[(False, 3)]:.1f})   split-K route "
          f"{out[(True, 0)]:.1f} (pdl {out[(True, 3)]:.1f})", flush=True)
    for name, fn in pieces.items():
        print(f"   {name:28s} {timed(fn):6.1f} us", flush=True)
