"""Shared pieces of the decode-stack experiment: constants of decode_stack.cu, host packing of the per-owner slabs and the
plain torch fp32 reference of the stack (used by run.py on the GPU and by emulate.py on the CPU)."""
import math

import torch

DM, NH, HD, FF, OWNERS, MAXL = 1024, 16, 64, 4096, 128, 16
NA, NC, NF, XS, D2S = 24, 8, 32, 1032, 36
bf = torch.bfloat16
EPS, SCALE = 1e-6, 1.0 / math.sqrt(HD)


def pack_rows(w, rows_per_owner):
    """[N, 1024] bf16 → [OWNERS][rows_per_owner][XS] (zero padded): the shared-memory image of each owner's slab."""
    assert w.shape[0] == OWNERS * rows_per_owner
    out = torch.zeros(OWNERS, rows_per_owner, XS, dtype=bf, device=w.device)
    out[..., :DM] = w.view(OWNERS, rows_per_owner, DM)
    return out.contiguous()


def pack_w2(w2):
    """W2 [1024, 4096] → [OWNERS][1024 n][D2S]: owner c holds the k-slice [32c, 32c + 32) of every output row."""
    out = torch.zeros(OWNERS, DM, D2S, dtype=bf, device=w2.device)
    out[..., :NF] = w2.view(DM, OWNERS, NF).permute(1, 0, 2)
    return out.contiguous()


def alibi_slopes(h, dev):
    return torch.tensor([2.0 ** (-(i + 1) / 2) for i in range(h)], dtype=torch.float32, device=dev)


def make_layers(L, dev, seed=0):
    torch.manual_seed(seed)
    g = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc)          # noqa: E731
    return [dict(w_in=g(3 * DM, DM, sc=DM ** -0.5).to(bf), w_out=g(DM, DM, sc=DM ** -0.5).to(bf),
                 w1=g(FF, DM, sc=DM ** -0.5).to(bf), w2=g(DM, FF, sc=FF ** -0.5).to(bf),
                 n1=1 + 0.1 * g(DM), n3=1 + 0.1 * g(DM), b1=0.1 * g(FF), b2=0.1 * g(DM)) for _ in range(L)]


def rms(x, w):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + EPS) * w


def reference(layers, slopes, x, kcs, vcs, pos):
    """fp32 stack on [B, 1024]; kcs / vcs: per-layer fp32 caches [B, H, Tmax, 64] holding `pos` past rows."""
    B = x.shape[0]
    for ly, kc, vc in zip(layers, kcs, vcs):
        qkv = rms(x, ly["n1"]) @ ly["w_in"].float().t()
        q, k, v = (t.view(B, NH, HD) for t in qkv.split(DM, -1))
        kc[:, :, pos] = k.to(bf).float()
        vc[:, :, pos] = v.to(bf).float()
        s = torch.einsum("bhd,bhtd->bht", q, kc[:, :, :pos + 1]) * SCALE
        s = s - slopes.view(1, NH, 1) * (pos - torch.arange(pos + 1, device=x.device)).view(1, 1, -1)
        o = torch.einsum("bht,bhtd->bhd", torch.softmax(s, -1), vc[:, :, :pos + 1]).reshape(B, DM)
        x = x + o.to(bf).float() @ ly["w_out"].float().t()
        hdn = torch.nn.functional.gelu(rms(x, ly["n3"]) @ ly["w1"].float().t() + ly["b1"])
        x = x + hdn @ ly["w2"].float().t() + ly["b2"]
    return x
