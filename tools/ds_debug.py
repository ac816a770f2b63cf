"""Stage-by-stage check of the persistent decode-step kernel (vg_decode_step) against a torch mirror of its dataflow.
usage: python tools/ds_debug.py [--small] [--batch B] [--prefix k1,k2,...] [--mode 0|1]
Runs only the first k phases of the step (DecodeStepEngine(debug_phases=k)) and compares the buffer that phase produced."""
import argparse
import copy
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vae_gslm_b200 import _lib
from vae_gslm_b200.decode_step import DecodeStepEngine
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR
from vae_gslm_b200.training_lib.trainer import init_weights

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--prompt", type=int, default=37)
ap.add_argument("--prefix", default="1,2,3,4,5,6,11,0")
ap.add_argument("--mode", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib.load()
torch.manual_seed(0)
if args.small:
    g = torch.load(os.path.join(ROOT, "tests", "golden", "lvtr_small.pt"), map_location="cpu", weights_only=False)
    model = LVTR(Hparams.from_dict(copy.deepcopy(g["config"])), input_dim=g["n_mels"])
    model.load_state_dict(g["state_dict"], strict=False)
    vocab = g["config"]["tokens"]["vocab_size"]
else:
    hp = Hparams.from_yamlfile(os.path.join(ROOT, "vae_gslm_b200", "configs", "train", "speech", "vae-gslm.yaml"))
    model = LVTR(hp.model, input_dim=80)
    model.apply(init_weights)
    vocab = 200
model = model.to(dev).set_compute_dtype(torch.bfloat16).eval()
model.use_decode_engine = False                   # prefill / reference steps on the layer-by-layer path
B, P = args.batch, args.prompt
stack = model.transformer[0]
stack.cache_len_hint = P + 40
prompt = torch.cat([torch.randint(0, vocab, (B, P, 1), device=dev).float(), torch.randn(B, P, 4, device=dev)], -1)
o = model.step(prompt, past_kv=None, temperature=0.0, push_init_state=True, greedy=True)
kv = o["kv"]
cache = kv[0].cache
pos = cache.length
state = o["output"][:, -1:]
ids, zin = state[..., 0].long(), state[..., 1:].float()
fuser = model.token_fuser.linear
u = (torch.nn.functional.embedding(ids, model.token_embedding.weight)
     + torch.relu(torch.nn.functional.linear(zin, fuser.weight, fuser.bias)))[:, 0]
u16 = u.to(torch.bfloat16)
cache_snapshot = cache.buf.clone()


def bf(t):
    return t.to(torch.bfloat16).float()


def mirror():
    """torch mirror of the kernel's dataflow (bf16 weights and X operands, fp32 accumulation and residual)"""
    out = {}
    W = lambda p: p.detach().to(torch.bfloat16).float()          # noqa: E731
    eps = stack.layers[0].norm1.eps
    d = stack.hp.layer.dim
    H = stack.hp.layer.self_attn.nheads
    h = u16.float() @ W(stack.linear.weight).t()
    out["in"] = h.clone()
    slopes = stack.rpe.slopes.float()
    for i, lyr in enumerate(stack.layers):
        rstd = torch.rsqrt((h * h).mean(-1, keepdim=True) + eps)
        acc = bf(h * lyr.norm1.scale.float()) @ W(lyr.self_attn.in_proj.weight).t()
        out[f"qkv{i}"] = acc.clone()
        qkv = acc * rstd
        q, k, v = qkv[:, :d], bf(qkv[:, d:2 * d]), bf(qkv[:, 2 * d:])
        kc = cache_snapshot[i, 0, :, :, :pos].float()             # [B,H,pos,64]
        vc = cache_snapshot[i, 1, :, :, :pos].float()
        kk = torch.cat([kc, k.view(B, H, 1, 64)], 2)
        vv = torch.cat([vc, v.view(B, H, 1, 64)], 2)
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


ref = mirror()


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


names = None
for k in [int(x) for x in args.prefix.split(",")]:
    cache.buf.copy_(cache_snapshot)
    cache.length = pos
    eng = DecodeStepEngine(model, B, dev, barrier_mode=args.mode, debug_phases=k)
    names = eng.phase_names
    try:
        eng.run(u16, kv)
        torch.cuda.synchronize()
    except Exception as e:                                    # noqa: BLE001
        print("FAILED", k, names[-1], e, [hex(int(x)) for x in eng.debug.cpu().tolist()])
        raise
    last = names[-1]
    d = stack.hp.layer.dim
    if last == "in" or last.startswith(("out", "ffn2_")):
        got, want = eng.h, ref[last]
    elif last.startswith("qkv"):
        got, want = eng.qkv_acc.view(B, -1), ref[last]
    elif last.startswith("attn"):
        i = int(last[4:])
        got, want = (eng.o if not eng.late_merge else None), ref[last]
        if got is None:
            continue
        print(f"   cache append k {rel(cache.buf[i, 0, :, :, pos].reshape(B, -1), ref[f'k{i}']):.2e} "
              f"v {rel(cache.buf[i, 1, :, :, pos].reshape(B, -1), ref[f'v{i}']):.2e}")
    elif last.startswith("ffn1_"):
        got, want = eng.f1_acc.view(B, -1), ref[last]
    elif last == "split":
        got, want = eng.cg_acc.view(B, -1), ref[last]
    else:
        got, want = eng.logits, ref["logits"]
        print(f"   H {rel(eng.H, ref['H']):.2e} head {rel(eng.head, ref['head']):.2e}")
    print(f"phases {len(names):3d} (last = {last:8s}): max rel err {rel(got, want):.3e}   "
          f"[R,S of last phase units: {eng.last_RS}]", flush=True)

# the whole step against the layer-by-layer product path
cache.buf.copy_(cache_snapshot)
cache.length = pos
o2 = model.step(state, past_kv=kv, temperature=0.0, greedy=True, return_logits=True)
print("layerwise path vs mirror: logits", rel(o2["logits"].view(B, -1), ref["logits"]), "H",
      rel(o2["transformer_latent"].value.view(B, -1), ref["H"]))
