"""per-tensor gradient errors of the fp32 parity mode at the full configuration (debug aid for tests/test_parity_full_gpu.py)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_parity_full_gpu as T
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B, Tn = int(sys.argv[1]), int(sys.argv[2])
model, cfg = T.full_model()
batch, rng = T.synthetic(B, Tn)
sd_src = dict(model.state_dict()); sd_src["_param_names"] = {n for n, _ in model.named_parameters()}
ref = T.oracle_run(sd_src, cfg, batch, rng, 0.04, autocast=False)
model.set_compute_dtype(torch.float32)
out = model(T.TensorMask(batch["x"], batch["mask"]), utterance=T.TensorMask(batch["utterance"], batch["utt_mask"]), **rng)
terms = T.assemble_loss(out, kld_weight=0.04)
terms["loss"].backward()
errs = sorted(((T.frob_rel(p.grad, ref["grads"][n]), n) for n, p in model.named_parameters()), reverse=True)
for e, n in errs[:12]:
    print(f"{e:.3e} {n}")
print("median", errs[len(errs) // 2][0])
n = "token_spliter.linear.bias"
g, r = dict(model.named_parameters())[n].grad, ref["grads"][n]
d = (g - r).abs()
print("bias grad: max abs diff", float(d.max()), "at", int(d.argmax()), "ref there", float(r[d.argmax()]), "ref norm", float(r.norm()), "n big", int((d > 1e-3 * r.abs().max()).sum()))
g, r = dict(model.named_parameters())["token_spliter.linear.weight"].grad, ref["grads"]["token_spliter.linear.weight"]
rowerr = (g - r).norm(dim=1) / (r.norm(dim=1) + 1e-20)
print("weight grad rows with rel err > 1e-3:", int((rowerr > 1e-3).sum()), "worst rows", rowerr.topk(5))
