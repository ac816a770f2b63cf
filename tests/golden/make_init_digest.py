"""Digest of the REAL reference's seeded random init at the full VAE-GSLM configuration.

Run in the build container only:  python tests/golden/make_init_digest.py
Builds ``models.speech.lvtr.LVTR`` from /root/reference/configs/train/speech/vae-gslm.yaml under
``torch.manual_seed(0)``, applies the rules of ``BaseTrainer.init_weights`` (training_lib/trainer.py:113-125; the
trainer class itself needs `lightning`, so its 12 lines are applied through a stand-in object carrying the same method)
and stores, per state-dict entry in order: name, shape, dtype and the SHA-256 of the raw bytes.
tests/test_host_logic.py::test_seeded_init_is_bit_identical_to_reference holds this repo's model to it.
"""
import hashlib
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_shim import REFERENCE_ROOT, import_reference  # noqa: E402


def digest_state_dict(sd) -> list:
    out = []
    for k, v in sd.items():
        t = v.detach().cpu().contiguous()
        out.append({"name": k, "shape": list(t.shape), "dtype": str(t.dtype).replace("torch.", ""),
                    "sha256": hashlib.sha256(t.numpy().tobytes()).hexdigest()})
    return out


def main():
    LVTR, Hparams, _ = import_reference()
    import ast
    import inspect
    import torch.nn as nn  # noqa: F401  (used by the reference method's globals)
    hp = Hparams.from_yamlfile(os.path.join(REFERENCE_ROOT, "configs", "train", "speech", "vae-gslm.yaml"))
    # BaseTrainer.init_weights, executed from the reference's own source text (the class cannot be imported: lightning)
    src = open(os.path.join(REFERENCE_ROOT, "training_lib", "trainer.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "init_weights")
    ns = {"nn": torch.nn}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref:init_weights", "exec"), ns)
    holder = types.SimpleNamespace(hp=hp)
    torch.manual_seed(0)
    model = LVTR(hp.model, input_dim=80)
    model.apply(lambda m: ns["init_weights"](holder, m))
    d = {"seed": 0, "config": "configs/train/speech/vae-gslm.yaml", "n_params": sum(p.numel() for p in model.parameters()),
         "entries": digest_state_dict(model.state_dict())}
    with open(os.path.join(HERE, "init_digest.json"), "w") as f:
        json.dump(d, f, indent=0)
    print("wrote", len(d["entries"]), "entries,", d["n_params"], "parameters")


if __name__ == "__main__":
    main()
