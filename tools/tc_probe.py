"""Stand-alone diagnostic for the tcgen05 GEMM (run on the GPU box under `timeout`): every operand layout,
with per-64x64-block error maps so a descriptor/layout mistake can be localised from one run."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def run(tag, M, N, K, ta, tb, out_dtype=torch.float32):
    a = torch.randn((K, M) if ta else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((N, K) if tb else (K, N), device=dev).to(torch.bfloat16)
    print(f"[{tag}] M={M} N={N} K={K} trans_a={ta} trans_b={tb} ...", flush=True)
    out = ops.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=out_dtype, backend=ops.GEMM_TCGEN05)
    torch.cuda.synchronize()
    A = a.float().t() if ta else a.float()
    B = b.float().t() if tb else b.float()
    ref = A @ B
    err = (out.float() - ref).abs()
    rel = float(err.max() / ref.abs().max())
    print(f"    rel err {rel:.3e}  {'OK' if rel < 5e-3 else 'MISMATCH'}", flush=True)
    if rel >= 5e-3:
        bm = (M + 63) // 64
        bn = (N + 63) // 64
        for i in range(min(bm, 4)):
            row = []
            for j in range(min(bn, 8)):
                blk = err[i * 64:(i + 1) * 64, j * 64:(j + 1) * 64]
                row.append(f"{float(blk.max()):8.2f}")
            print("    block max err:", " ".join(row))
        print("    out[0,:8]", out[0, :8].float().tolist())
        print("    ref[0,:8]", ref[0, :8].tolist())
    return rel < 5e-3


ok = True
for (M, N, K) in [(128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 512, 128), (304, 200, 72)]:
    for ta, tb in [(False, True), (False, False), (True, True), (True, False)]:
        ok &= run("probe", M, N, K, ta, tb)
ok &= run("big", 5120, 4096, 1024, False, True, torch.bfloat16)
ok &= run("wgrad", 4096, 1024, 5120, True, False)
print("ALL OK" if ok else "SOME MISMATCH")
# quick timing of the forward FFN1 shape
import time
a = torch.randn(5120, 1024, device=dev).to(torch.bfloat16)
w = torch.randn(4096, 1024, device=dev).to(torch.bfloat16)
for be, name in ((ops.GEMM_TCGEN05, "tcgen05"),):
    for _ in range(3):
        ops.gemm(a, w, backend=be)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm(a, w, backend=be)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: 5120x4096x1024 {ms*1e3:.1f} us  {2*5120*4096*1024/ms/1e9:.1f} TFLOP/s")
for _ in range(3):
    a @ w.t()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    a @ w.t()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"cublas : 5120x4096x1024 {ms*1e3:.1f} us  {2*5120*4096*1024/ms/1e9:.1f} TFLOP/s")
