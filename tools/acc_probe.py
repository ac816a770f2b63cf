import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_model_gpu as T
from vae_gslm_b200.trainers.speech.lvtr import assemble_loss
from vae_gslm_b200.utils.tensormask import TensorMask
golden = torch.load("/root/repo/tests/golden/lvtr_small.pt", map_location="cpu", weights_only=False)
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
model, arena, step, batch = T._train_setup(golden, lr=0.0, use_cuda_graph=False, accumulate=2, overlap_grads=False)
model.overlap_decoder = False
b2 = {k: v.clone() for k, v in batch.items()}
b2["x"] = batch["x"].flip(0).contiguous(); b2["mask"] = batch["mask"].flip(0).contiguous()
step.load(batch, 0); step.load(b2, 1)
torch.cuda.synchronize()
for k in b2:
    print(k, "slot1 == b2:", torch.equal(step.statics[1][k], b2[k]), "slot0 == batch:", torch.equal(step.statics[0][k], batch[k]))
def fwd(s):
    out = model(TensorMask(s["x"], s["mask"]), utterance=TensorMask(s["utterance"], s["utt_mask"]))
    return float(assemble_loss(out, kld_weight=0.04)["loss"].detach())
print("direct fwd slot0", fwd(step.statics[0]), "slot1", fwd(step.statics[1]), "b2 itself", fwd(b2))
print("again slot1", fwd(step.statics[1]), "slot0", fwd(step.statics[0]))
with torch.no_grad():
    print("no_grad slot1", fwd(step.statics[1]))
print("acc", float(step(lr=0.0)))
print("after acc: direct fwd slot1", fwd(step.statics[1]))
