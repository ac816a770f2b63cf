import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_model_gpu as T
from vae_gslm_b200.arena import ParamArena
from vae_gslm_b200.dp import GradReducer
from vae_gslm_b200.trainers.speech.lvtr import TrainStep
golden = torch.load(os.path.join(os.path.dirname(T.__file__), "golden", "lvtr_small.pt"), weights_only=False)
DEV = "cuda"
i = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in golden["inputs"].items()}
batch = {k: i[k] for k in ("x", "mask", "utterance", "utt_mask")}
draws = {k: i[k] for k in ("eps_q", "init_state", "eps_p", "diff_t", "diff_noise")}
model = T.build_small(golden, torch.bfloat16)
inner = model.forward
model.forward = lambda x, _f=inner, **kw: _f(x, **kw, **draws)
arena = ParamArena(model, weight_decay=0.1)
arena._BIG = 1 << 12
step = TrainStep(model, arena, GradReducer(arena), batch, lr=0.0, kld_weight=0.04, use_cuda_graph=False)
print("direct_calls", len(arena._direct_calls), "autograd_hits", len(arena._autograd_hits), "overwrite", len(arena._overwrite),
      "calibrating", arena._calibrating, "segments", [None if s is None else s[2] for s in (arena._segments or [])])
names = {id(p): n for g in arena.groups for n, p in zip(g.names, g.params)}
both = [names[k] for k in arena._direct_calls if arena._autograd_hits.get(k, 0)]
print("direct AND autograd:", len(both), both[:8])
print("sizes of overwrite:", sorted((p.numel() for g in arena.groups for p in g.params if id(p) in arena._overwrite), reverse=True)[:10])
