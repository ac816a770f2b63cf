timeout 300 python -m pytest tests/test_model_gpu.py -q --tb=short -k "graphed_ddim" 2>&1 | tail -15
timeout 600 python - <<'PY'
import torch, bench
from vae_gslm_b200 import _lib
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR
from vae_gslm_b200.training_lib.trainer import init_weights
dev=torch.device("cuda",0); _lib.load(); torch.manual_seed(0)
hp=Hparams.from_yamlfile(bench.CFG); model=LVTR(hp.model, input_dim=bench.N_MELS); model.apply(init_weights)
model=model.to(dev).set_compute_dtype(torch.bfloat16).eval()
print(bench.ddim_bench(model, dev))
PY
