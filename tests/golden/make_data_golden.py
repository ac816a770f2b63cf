"""Generate tests/golden/data_pipeline.pt: items and the collated batch of the REAL reference ``DiscreteTokenDataset``
(/root/reference/data/dataset.py) on the synthetic corpus of tests/test_data_pipeline.py (seeded), so that the parity
check also runs where the reference tree is absent.  Run in the build container: ``python tests/golden/make_data_golden.py``."""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, here)
sys.path.insert(0, os.path.dirname(here))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from ref_shim import import_reference  # noqa: E402
from test_data_pipeline import _corpus  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    cfg, mel, hubert, rescale = _corpus(Path(d))
    _, RefHp, _ = import_reference()
    from data.dataset import DiscreteTokenDataset
    mk = lambda x: RefHp.from_json(json.dumps(x)) if hasattr(RefHp, "from_json") else RefHp.from_dict(x)
    ds = DiscreteTokenDataset(mk(cfg), mk(mel), mk(hubert), mk(rescale))
    torch.manual_seed(123)
    np.random.seed(123)
    items = [ds[i] for i in range(len(ds))]
    batch = ds.seqCollate(items)
    out = {"audios": [os.path.relpath(a, cfg["wavdir"]) for a in ds.audios], "items": items,
           "batch": {k: (v.value, v.mask) for k, v in batch.items()}}
    torch.save(out, os.path.join(here, "data_pipeline.pt"))
    print({k: tuple(v[0].shape) for k, v in out["batch"].items()}, len(items))
