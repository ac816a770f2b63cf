"""cached-generation throughput (BASELINE configs[2]): 3 s prompt, single-frame steps replayed from a CUDA graph,
batch sweep; frames/s against the HBM roofline of SURVEY §8d.  usage: python tools/decode_bench.py [B ...]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

from vae_gslm_b200 import _lib
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.speech.lvtr import LVTR
from vae_gslm_b200.training_lib.trainer import init_weights

dev = torch.device("cuda", 0)
_lib.load()
torch.manual_seed(0)
hp = Hparams.from_yamlfile(bench.CFG)
model = LVTR(hp.model, input_dim=bench.N_MELS)
model.apply(init_weights)
model = model.to(dev).set_compute_dtype(torch.bfloat16).eval()
batches = tuple(int(a) for a in sys.argv[1:] if a.isdigit()) or (1, 8, 64, 256)
if "--force-engine" in sys.argv:
    model.decode_engine_max_batch = 256
for engine in ((True, False) if "--both" in sys.argv else (True,)):
    model.use_decode_engine = engine
    out = bench.decode_bench(model, dev, bench.measured_peaks(), batches=batches)
    print("engine" if engine else "layerwise", json.dumps(out), flush=True)
