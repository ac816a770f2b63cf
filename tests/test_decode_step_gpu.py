"""vg_decode_step (the persistent single-launch generation step, csrc/decode_step.cu + decode_step.py) against a plain
torch restatement of the same arithmetic — bf16 weights and GEMM operands, fp32 accumulation and residual stream, the
reference's RMSNorm / attention-with-ALiBi / GELU FFN (modules/transformer/layers.py:41-93,134-195, attention.py:52-85) —
on the same cache state: every intermediate buffer of the first layer (phase prefixes) and the outputs of the whole step,
for the unit decompositions of batch 1 ... 256 (1 / 2 / 4 row copies, one and two M tiles, warp / CTA attention items)."""
import copy
import os

import pytest
import torch

from vae_gslm_b200.decode_step import DecodeStepEngine                  # noqa: E402
from vae_gslm_b200.hparams.hp import Hparams                            # noqa: E402
from vae_gslm_b200.models.speech.lvtr import LVTR                       # noqa: E402

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bf(t):
    return t.to(torch.bfloat16).float()


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


def mirror_step(model, u16, cache_buf, pos):
    """torch restatement of one cached step: returns every buffer the kernel's phases produce"""
    stack = model.transformer[0]
    out = {}
    W = lambda p: p.detach().to(torch.bfloat16).float()          # noqa: E731
    eps = stack.layers[0].norm1.eps
    d = stack.hp.layer.dim
    H = stack.hp.layer.self_attn.nheads
    B = u16.shape[0]
    dev = u16.device
    h = u16.float() @ W(stack.linear.weight).t()
    out["in"] = h.clone()
    slopes = stack.rpe.slopes.float()
    for i, lyr in enumerate(stack.layers):
        rstd = torch.rsqrt((h * h).mean(-1, keepdim=True) + eps)
        acc = bf(h * lyr.norm1.scale.float()) @ W(lyr.self_attn.in_proj.weight).t()
        out[f"qkv{i}"] = acc.clone()
        qkv = acc * rstd
        q, k, v = qkv[:, :d], bf(qkv[:, d:2 * d]), bf(qkv[:, 2 * d:])
        kk = torch.cat([cache_buf[i, 0, :, :, :pos].float(), k.view(B, H, 1, 64)], 2)
        vv = torch.cat([cache_buf[i, 1, :, :, :pos].float(), v.view(B, H, 1, 64)], 2)
        s = torch.einsum("bhd,bhjd->bhj", q.view(B, H, 64), kk) / 8.0
        j = torch.arange(pos + 1, device=dev)
        s = s - slopes.view(1, H, 1) * (pos - j).view(1, 1, -1)
        att = torch.einsum("bhj,bhjd->bhd", torch.softmax(s, -1), vv).reshape(B, d)
        out[f"attn{i}"] = bf(att)
        out[f"k{i}"], out[f"v{i}"] = k, v
        h = h + bf(att) @ W(lyr.self_attn.out_proj.weight).t()
        out[f"out{i}"] = h.clone()
        rstd = torch.rsqrt((h * h).mean(-1, keepdim=True) + eps)
        f1 = bf(h * lyr.norm3.scale.float()) @ W(lyr.linear1.weight).t()
        out[f"ffn1_{i}"] = f1.clone()
        gl = bf(torch.nn.functional.gelu(f1 * rstd + lyr.linear1.bias.float()))
        h = h + gl @ W(lyr.linear2.weight).t() + lyr.linear2.bias.float()
        out[f"ffn2_{i}"] = h.clone()
    rstd = torch.rsqrt((h * h).mean(-1, keepdim=True) + eps)
    fn = stack.final_norm.scale.float()
    out["H"] = bf(h * fn * rstd)
    w_split, b_split = model._split_weights()
    cgacc = bf(h * fn) @ W(w_split).t()
    out["split"] = cgacc.clone()
    cg = bf(torch.relu(cgacc * rstd + b_split.float()))
    w_head, b_head = model._head_weights()
    out["head"] = cg[:, :d] @ W(w_head).t() + b_head.float()
    tp = model.token_predictor.linear
    out["logits"] = cg[:, d:] @ W(tp.weight).t() + tp.bias.float()
    return out


def phase_buffer(eng, last, B):
    if last == "in" or last.startswith(("out", "ffn2_")):
        return eng.h
    if last.startswith("qkv"):
        return eng.qkv_acc.view(B, -1)
    if last.startswith("attn"):
        if not eng.late_merge:
            return eng.o
        # late merge: the phase publishes (m, l, o[64]) per kv-split (exp2 domain); merge them as the out-projection does
        H = eng.nheads
        pt = eng.attn_partial.view(B * H, eng.nsplit, 72)
        m, l, o = pt[..., 0], pt[..., 1], pt[..., 8:]
        w = torch.exp2(m - m.max(-1, keepdim=True).values)
        w = torch.where(torch.isfinite(m), w, torch.zeros_like(w))
        att = (w[..., None] * o).sum(1) / (w * l).sum(1, keepdim=True)
        return bf(att.view(B, H * 64))
    if last.startswith("ffn1_"):
        return eng.f1_acc.view(B, -1)
    if last == "split":
        return eng.cg_acc.view(B, -1)
    return eng.logits


def build(small, golden):
    from vae_gslm_b200.training_lib.trainer import init_weights
    torch.manual_seed(0)
    if small:
        model = LVTR(Hparams.from_dict(copy.deepcopy(golden["config"])), input_dim=golden["n_mels"])
        model.load_state_dict(golden["state_dict"], strict=False)
        vocab = golden["config"]["tokens"]["vocab_size"]
    else:
        hp = Hparams.from_yamlfile(os.path.join(ROOT, "vae_gslm_b200", "configs", "train", "speech", "vae-gslm.yaml"))
        model = LVTR(hp.model, input_dim=80)
        model.apply(init_weights)
        g = torch.Generator().manual_seed(99)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith(".bias"):
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
        vocab = 200
    model = model.to(DEV).set_compute_dtype(torch.bfloat16).eval()
    return model, vocab


def prefill(model, vocab, B, P):
    model.use_decode_engine = False                       # prompt through the layer-by-layer path
    model.transformer[0].cache_len_hint = P + 40
    prompt = torch.cat([torch.randint(0, vocab, (B, P, 1), device=DEV).float(), torch.randn(B, P, 4, device=DEV)], -1)
    o = model.step(prompt, past_kv=None, temperature=0.0, push_init_state=True, greedy=True)
    kv = o["kv"]
    state = o["output"][:, -1:]
    fuser = model.token_fuser.linear
    u = (torch.nn.functional.embedding(state[..., 0].long(), model.token_embedding.weight)
         + torch.relu(torch.nn.functional.linear(state[..., 1:].float(), fuser.weight, fuser.bias)))[:, 0]
    return kv, u.to(torch.bfloat16)


@pytest.mark.gpu
@pytest.mark.parametrize("small,B,P,prefixes", [
    (True, 3, 37, (1, 2, 3, 4, 5, 6, 0)),
    (False, 1, 37, (2, 3, 4, 5, 6, 0)),
    (False, 5, 301, (3, 0)),
    (False, 16, 70, (3, 0)),            # warp-per-item attention with kv-splits
    (False, 40, 37, (2, 0)),            # two row copies
    (False, 130, 37, (0,)),             # two M tiles
    (False, 256, 150, (2, 3, 0)),
])
@pytest.mark.parametrize("barrier_mode", [1, 2])
def test_decode_step_kernel_against_torch(golden, small, B, P, prefixes, barrier_mode):
    if barrier_mode == 2 and (B not in (1, 40)):
        pytest.skip("the flag barrier is exercised at two batch sizes")
    model, vocab = build(small, golden)
    kv, u16 = prefill(model, vocab, B, P)
    cache = kv[0].cache
    pos = cache.length
    snapshot = cache.buf.clone()
    ref = mirror_step(model, u16, snapshot, pos)
    for k in prefixes:
        cache.buf.copy_(snapshot)
        cache.length = pos
        eng = DecodeStepEngine(model, B, DEV, barrier_mode=barrier_mode, debug_phases=k)
        for rep in range(2):                              # twice: the kernel leaves its barrier / ticket state ready to reuse
            cache.buf.copy_(snapshot)
            cache.length = pos
            eng.run(u16, kv)
        torch.cuda.synchronize()
        last = eng.phase_names[-1]
        want = ref["logits"] if last == "heads" else ref[last]
        # the stack-input linear: summation order only; from the first RMSNorm on the GEMM operand is a bf16 ROUNDING of an
        # fp32 value that differs in its last bits between the two sides (a flipped element moves a sum by ~1e-4 of the row
        # maximum); the attention output is bf16 itself (one ulp at the row maximum = 3.9e-3); the whole step: 16 layers
        tol = 1e-5 if last == "in" else (8e-3 if last.startswith("attn") else (3e-3 if k else 1.5e-2))
        err = rel(phase_buffer(eng, last, B), want)
        assert err < tol, (last, err)
        if last.startswith("attn"):
            i = int(last[4:])
            assert rel(cache.buf[i, 0, :, :, pos].reshape(B, -1), ref[f"k{i}"]) < 8e-3      # one bf16 ulp
            assert rel(cache.buf[i, 1, :, :, pos].reshape(B, -1), ref[f"v{i}"]) < 8e-3
        if last == "heads":
            assert rel(eng.H, ref["H"]) < 1.5e-2 and rel(eng.head, ref["head"]) < 1.5e-2
            assert cache.length == pos + 1
    # the position / epoch carried on the device (CUDA-graph replay mode)
    assert int(eng.epoch) == 2 * (eng.NP - 1)


def test_decode_step_planner_contract():
    """host side of the kernel's contracts: units cover every output feature exactly once per k-slice, k-blocks per unit
    are powers of two, streams are 1 KB aligned, and the packed bytes are the SWIZZLE_128B K-major image of the weights."""
    from vae_gslm_b200.decode_step import pack_units, split_factor
    for (N, K) in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096), (2048, 1024), (520, 1024), (200, 1024), (384, 128)):
        for rmax in (128, 256):
            R, S = split_factor(N, K, rmax=rmax)
            assert R % 16 == 0 and R <= rmax and (K // 64) % S == 0
            nkb = K // 64 // S
            assert nkb & (nkb - 1) == 0
    W = torch.randn(48, 256).to(torch.bfloat16)
    R, S = 16, 2
    P = pack_units(W, R, S)
    nkb = 256 // 64 // S
    for slab in range(3):
        for s in range(S):
            u = P[slab, s].view(-1)
            for kb in range(nkb):
                for row in range(R):
                    for k in range(0, 64, 8):
                        addr = (kb * R * 128 + (row // 8) * 1024 + (row % 8) * 128 + (((k // 8) ^ (row % 8)) * 16)) // 2
                        assert torch.equal(u[addr:addr + 8], W[slab * R + row, s * nkb * 64 + kb * 64 + k: s * nkb * 64 + kb * 64 + k + 8])


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 16])
def test_decode_step_kernel_ticket_merge_route(golden, B):
    """the in-phase merge of the kv-split partials (ticket + last arriver; late_merge=False) stays correct: one CTA per
    item (B = 1, 4 splits) and one warp per item (B = 16, 4 splits)"""
    model, vocab = build(False, golden)
    kv, u16 = prefill(model, vocab, B, 70)
    cache = kv[0].cache
    pos = cache.length
    snapshot = cache.buf.clone()
    ref = mirror_step(model, u16, snapshot, pos)
    for k in (3, 0):
        cache.buf.copy_(snapshot)
        cache.length = pos
        eng = DecodeStepEngine(model, B, DEV, debug_phases=k, late_merge=False)
        assert eng.nsplit > 1 and not eng.late_merge
        for rep in range(2):
            cache.buf.copy_(snapshot)
            cache.length = pos
            eng.run(u16, kv)
        torch.cuda.synchronize()
        last = eng.phase_names[-1]
        want = ref["logits"] if last == "heads" else ref[last]
        assert rel(phase_buffer(eng, last, B), want) < (8e-3 if k else 1.5e-2)
        assert int(eng.tickets.abs().sum()) == 0
