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
kinds = [a.split("=")[1] for a in sys.argv if a.startswith("--kind=")] or ["step"]
if "--both" in sys.argv:
    kinds = ["step", "linear", "layerwise"]
gen = int(([a.split("=")[1] for a in sys.argv if a.startswith("--gen=")] or ["500"])[0])
for kind in kinds:
    model.use_decode_engine = kind != "layerwise"
    model.decode_engine_kind = kind
    if kind == "linear" and "--force-engine" not in sys.argv:
        model.decode_engine_max_batch = 96
    out = bench.decode_bench(model, dev, bench.measured_peaks(), batches=batches, gen=gen, ddim=False)
    for k, v in out.items():
        print(kind, k, json.dumps(v), flush=True)
