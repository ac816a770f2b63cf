"""prefill + a few eager single-frame steps at batch B — target of ncu launch lists for the decode path."""
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
dev = torch.device("cuda", 0)
_lib.load()
torch.manual_seed(0)
hp = Hparams.from_yamlfile(bench.CFG)
model = LVTR(hp.model, input_dim=bench.N_MELS)
model.apply(init_weights)
model = model.to(dev).set_compute_dtype(torch.bfloat16).eval()
g = torch.Generator().manual_seed(7)
prior = torch.cat([torch.randint(0, 200, (B, 150, 1), generator=g).float(), torch.randn(B, 150, 4, generator=g)], -1).to(dev)
model.transformer[0].cache_len_hint = 150 + 1 + 32
o = model.step(prior, past_kv=None, temperature=0.85, token_temperature=0.85, push_init_state=True)
state, kv = o["output"][:, -1:], o["kv"]
for _ in range(4):
    o = model.step(state, past_kv=kv, temperature=0.85, token_temperature=0.85)
    state, kv = o["output"], o["kv"]
torch.cuda.synchronize()
