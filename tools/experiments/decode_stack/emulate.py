"""CPU emulation of decode_stack.cu's DATAFLOW (no GPU): the same host-packed slabs, the same owner partition of every
GEMM, FFN2 as per-owner K-slice partial sums, and the double-buffered residual / accumulator scheme of the design notes —
executed phase by phase with torch on the CPU and compared with the plain reference.  It validates the packing and the
index arithmetic the kernel relies on (which feature lives in which owner's slab row, the [n][k] layout of the FFN2
slice, which buffer is read / written / cleared in which phase), not the CUDA code itself.
Run: python tools/experiments/decode_stack/emulate.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import (DM, EPS, FF, HD, NA, NC, NF, NH, OWNERS, SCALE, alibi_slopes, bf, make_layers, pack_rows, pack_w2,  # noqa: E402
                    reference)

dev = torch.device("cpu")
L, B, POS, TMAX = 3, 5, 37, 48
layers = make_layers(L, dev)
slopes = alibi_slopes(NH, dev)
packed = [dict(A=pack_rows(ly["w_in"], NA), C=pack_rows(ly["w_out"], NC), D1=pack_rows(ly["w1"], NF), D2=pack_w2(ly["w2"]))
          for ly in layers]
torch.manual_seed(1)
x0 = torch.randn(B, DM)
kc = [(0.5 * torch.randn(B, NH, TMAX, HD)).to(bf) for _ in range(L)]
vc = [(0.5 * torch.randn(B, NH, TMAX, HD)).to(bf) for _ in range(L)]
want = reference(layers, slopes, x0.clone(), [k.float() for k in kc], [v.float() for v in vc], POS)

hres = [x0.clone(), torch.zeros(B, DM)]
facc = [torch.zeros(B, DM), torch.zeros(B, DM)]
qbuf, abuf = torch.zeros(B, DM), torch.zeros(B, DM, dtype=bf)
log2e = 1.4426950408889634


def norm_tile(x, w):
    return (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + EPS) * w).to(bf).float()     # xa is bf16 in the kernel


for l, (ly, pk) in enumerate(zip(layers, packed)):
    p = l & 1
    # ---- phase A (every owner forms the same x; owner c writes its 8 merged columns and clears its accumulator columns)
    x = hres[p] + (facc[p] + layers[l - 1]["b2"] if l > 0 else 0)
    xa = norm_tile(x, ly["n1"])
    new_h, new_f = hres[p ^ 1].clone(), facc[p ^ 1].clone()
    for c in range(OWNERS):
        new_h[:, c * NC:(c + 1) * NC] = x[:, c * NC:(c + 1) * NC]
        new_f[:, c * NC:(c + 1) * NC] = 0
        out = xa @ pk["A"][c][:, :DM].float().t()                      # [B, 24]
        for n in range(NA):
            f = c * NA + n
            which, col = f // DM, f % DM
            h, dd = col // HD, col % HD
            if which == 0:
                qbuf[:, col] = out[:, n] * SCALE * log2e
            else:
                (kc if which == 1 else vc)[l][:, h, POS, dd] = out[:, n].to(bf)
    hres[p ^ 1], facc[p ^ 1] = new_h, new_f
    # ---- phase B (item = (b, head); scores in the log2 domain)
    for b in range(B):
        for h in range(NH):
            q = qbuf[b, h * HD:(h + 1) * HD]
            K, V = kc[l][b, h, :POS + 1].float(), vc[l][b, h, :POS + 1].float()
            s = K @ q - slopes[h] * log2e * (POS - torch.arange(POS + 1)).float()
            pr = torch.exp2(s - s.max())
            abuf[b, h * HD:(h + 1) * HD] = ((pr @ V) / pr.sum()).to(bf)
    # ---- phase C (owner adds its 8 out-projection features into the residual)
    xa = abuf.float()
    for c in range(OWNERS):
        hres[p ^ 1][:, c * NC:(c + 1) * NC] += xa @ pk["C"][c][:, :DM].float().t()
    # ---- phase D (FFN1 slice + GELU, then FFN2 partial sums from the [n][k] slice of W2)
    xa = norm_tile(hres[p ^ 1], ly["n3"])
    for c in range(OWNERS):
        hid = xa @ pk["D1"][c][:, :DM].float().t() + ly["b1"][c * NF:(c + 1) * NF]
        g = torch.nn.functional.gelu(hid).to(bf).float()               # the hidden tile is bf16 in shared memory
        facc[p ^ 1] += g @ pk["D2"][c][:, :NF].float().t()             # [B, 32] x [32, 1024]
got = hres[L & 1] + facc[L & 1] + layers[L - 1]["b2"]
err = float((got - want).abs().max() / want.abs().max())
print(f"emulated dataflow vs reference: max-norm relative error {err:.3e}")
assert err < 2e-2, err
print("OK: packing, owner partition, FFN2 K-slices and the double-buffered residual scheme reproduce the reference")
