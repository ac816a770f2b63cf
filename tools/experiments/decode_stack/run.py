"""Check and time the persistent decode-stack kernel (decode_stack.cu) on a B200 against a plain torch fp32 reference of the
same 16-layer pre-LN stack (RMSNorm → QKV → cached causal ALiBi attention → out-proj → RMSNorm → FFN-GELU), random weights.
Build first: `make -C tools/experiments/decode_stack`.  Usage: python tools/experiments/decode_stack/run.py [L] [prompt_len]"""
import ctypes as C
import math
import os
import sys

import torch

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, here)
lib = C.CDLL(os.path.join(here, "libds.so"))
from common import (D2S, DM, EPS, FF, HD, MAXL, NA, NC, NF, NH, OWNERS, SCALE, XS, alibi_slopes, bf, make_layers, pack_rows,  # noqa: E402,F401
                    pack_w2, reference)

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
PROMPT = int(sys.argv[2]) if len(sys.argv) > 2 else 180
dev = torch.device("cuda")


class Layer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("slabA", "slabC", "slabD1", "slabD2", "n1", "n3", "b1", "b2", "kc", "vc")]


class Params(C.Structure):
    _fields_ = [("layer", Layer * MAXL), ("L", C.c_int), ("B", C.c_int), ("pos", C.c_int), ("Tmax", C.c_int),
                ("eps", C.c_float), ("scale", C.c_float), ("slopes", C.c_void_p), ("hres", C.c_void_p * 2),
                ("facc", C.c_void_p * 2), ("qbuf", C.c_void_p), ("abuf", C.c_void_p), ("barrier", C.c_void_p),
                ("out", C.c_void_p)]


layers = make_layers(L, dev)
slopes = alibi_slopes(NH, dev)
eps, scale = EPS, SCALE
g = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc)          # noqa: E731
packed = [dict(A=pack_rows(ly["w_in"], NA), Cc=pack_rows(ly["w_out"], NC), D1=pack_rows(ly["w1"], NF), D2=pack_w2(ly["w2"]))
          for ly in layers]
assert packed[0]["A"][0].numel() * 2 == lib.ds_slab_bytes(0) and packed[0]["D2"][0].numel() * 2 == lib.ds_slab_bytes(3)
print(f"smem {lib.ds_smem_bytes()} B, max keys {lib.ds_max_keys()}, layers {L}, prompt {PROMPT}")
Tmax = PROMPT + 64
for B in (1, 4, 8):
    x0 = g(B, DM)
    kc = [g(B, NH, Tmax, HD, sc=0.5).to(bf) for _ in range(L)]
    vc = [g(B, NH, Tmax, HD, sc=0.5).to(bf) for _ in range(L)]
    want = reference(layers, slopes, x0.clone(), [k.float() for k in kc], [v.float() for v in vc], PROMPT)
    hres = [x0.clone(), torch.zeros_like(x0)]
    facc = [torch.zeros_like(x0), torch.zeros_like(x0)]
    qbuf, abuf = torch.zeros_like(x0), torch.zeros(B, DM, dtype=bf, device=dev)
    barrier, out = torch.zeros(4, dtype=torch.int32, device=dev), torch.zeros_like(x0)
    P = Params()
    for i, (ly, pk) in enumerate(zip(layers, packed)):
        for name, t in (("slabA", pk["A"]), ("slabC", pk["Cc"]), ("slabD1", pk["D1"]), ("slabD2", pk["D2"]), ("n1", ly["n1"]),
                        ("n3", ly["n3"]), ("b1", ly["b1"]), ("b2", ly["b2"]), ("kc", kc[i]), ("vc", vc[i])):
            setattr(P.layer[i], name, t.data_ptr())
    P.L, P.B, P.pos, P.Tmax, P.eps, P.scale = L, B, PROMPT, Tmax, eps, scale
    P.slopes = slopes.data_ptr()
    P.hres[0], P.hres[1], P.facc[0], P.facc[1] = hres[0].data_ptr(), hres[1].data_ptr(), facc[0].data_ptr(), facc[1].data_ptr()
    P.qbuf, P.abuf, P.barrier, P.out = qbuf.data_ptr(), abuf.data_ptr(), barrier.data_ptr(), out.data_ptr()
    rc = lib.ds_launch(C.byref(P), None)
    torch.cuda.synchronize()
    assert rc > 0, rc
    err = float((out - want).abs().max() / want.abs().max())
    print(f"B={B}: grid {rc}, max-norm relative error vs fp32 reference {err:.3e}  ({'OK' if err < 3e-2 else 'MISMATCH'})")
    # timing: the kernel re-runs on its own output state (weights and cache traffic are what is being timed)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    for rep in range(2):
        e0.record()
        for _ in range(n):
            for f in facc:
                f.zero_()
            lib.ds_launch(C.byref(P), None)
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    bytes_step = L * (3 * DM * DM + DM * DM + 2 * FF * DM) * 2 + B * L * 2 * NH * (PROMPT + 1) * HD * 2
    print(f"      {us:8.1f} us per stack pass; {bytes_step / 1e6:.1f} MB → {bytes_step / us / 1e3:.0f} GB/s "
          f"({bytes_step / us / 1e3 / 6547.5 * 100:.1f} % of the measured 6547.5 GB/s)")
