"""debug helper: bf16 cached decode vs the golden, with kernel backends toggled, plus tiny-M tcgen05 GEMMs."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib as L, ops
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR

dev = "cuda"
torch.manual_seed(0)
for M in (1, 2, 3, 7, 14, 48):
    for (N, K) in ((384, 128), (128, 64), (256, 128)):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16)
        out = ops.gemm(a, w, out_dtype=torch.float32, backend=ops.GEMM_TCGEN05)
        ref = a.float() @ w.float().t()
        print(f"gemm M={M} N={N} K={K} rel {float((out-ref).abs().max()/ref.abs().max()):.2e}")

g = torch.load("tests/golden/lvtr_small.pt", map_location="cpu", weights_only=False)
d = g["decode"]
for gemm_be, attn_be in ((ops.GEMM_SIMT, "simt"), (ops.GEMM_AUTO, "simt"), (ops.GEMM_SIMT, "auto"), (ops.GEMM_AUTO, "auto")):
    ops.GEMM_BACKEND = gemm_be
    L.set_attention_backend(attn_be)
    model = LVTR(Hparams.from_dict(copy.deepcopy(g["config"])), input_dim=g["n_mels"])
    model.load_state_dict(g["state_dict"], strict=False)
    model = model.to(dev).set_compute_dtype(torch.bfloat16).eval()
    state, kv = d["prompt"].to(dev), None
    errs = []
    for i, eps in enumerate(d["eps"]):
        o = model.step(state, past_kv=kv, temperature=0.85, push_init_state=(i == 0), eps=eps.to(dev), greedy=True,
                       init_state=d["init_state"].to(dev), return_logits=True)
        lat = o["transformer_latent"].value.float().cpu()
        e_lat = float((lat - d["latents"][i]).abs().max() / d["latents"][i].abs().max())
        lg = o["logits"].float().cpu()
        e_lg = float((lg - d["logits"][i]).abs().max() / d["logits"][i].abs().max())
        errs.append((round(e_lat, 4), round(e_lg, 4)))
        kv, state = o["kv"], d["outputs"][i][:, -1:].to(dev)
    print("gemm", "simt" if gemm_be == ops.GEMM_SIMT else "auto", "attn", attn_be, errs)
