"""prefill + single-frame steps at batch B — target of ncu launch lists for the decode path.  The LAST step is bracketed by
two torch.cuda._sleep marker kernels (spin_kernel) so that tools/decode_launch_summary.py can cut it out.
usage: python tools/decode_profile.py B [step|linear|layerwise] [prompt frames]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vae_gslm_b200 import _lib
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR
from vae_gslm_b200.training_lib.trainer import init_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "step"
P = int(sys.argv[3]) if len(sys.argv) > 3 else 400
dev = torch.device("cuda", 0)
_lib.load()
torch.manual_seed(0)
hp = Hparams.from_yamlfile(bench.CFG)
model = LVTR(hp.model, input_dim=bench.N_MELS)
model.apply(init_weights)
model = model.to(dev).set_compute_dtype(torch.bfloat16).eval()
model.use_decode_engine = kind != "layerwise"
model.decode_engine_kind = kind
model.decode_engine_max_batch = 256
g = torch.Generator().manual_seed(7)
prior = torch.cat([torch.randint(0, 200, (B, P, 1), generator=g).float(), torch.randn(B, P, 4, generator=g)], -1).to(dev)
model.transformer[0].cache_len_hint = P + 1 + 32
o = model.step(prior, past_kv=None, temperature=0.85, token_temperature=0.85, push_init_state=True)
state, kv = o["output"][:, -1:], o["kv"]
for i in range(4):
    if i == 3:
        torch.cuda._sleep(1000)
    o = model.step(state, past_kv=kv, temperature=0.85, token_temperature=0.85)
    state, kv = o["output"], o["kv"]
torch.cuda._sleep(1000)
torch.cuda.synchronize()
