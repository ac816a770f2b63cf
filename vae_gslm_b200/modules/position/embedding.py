"""Positional-encoding factory (reference ``modules/position/embedding.py:9-40``).  Only the encodings
on the VAE-GSLM path are built: ALiBi (transformer) and SinCos (diffusion time embedding); T5RPE and
Rotary are not reachable from the shipped configs and raise."""
from typing import Optional

from ...hparams.hp import Hparams
from .absolute import SinCos
from .alibi import ALiBi


def get_positional_encoding(name: str, hp: Hparams, ndim: Optional[int] = None, nheads: Optional[int] = None):
    if name == "SinCos":
        assert ndim is not None
        return SinCos(ndim, hp.get("maxpos", 10000), hp.get("fixed_pos", False), hp.get("scaled", False))
    if name == "ALiBi":
        assert nheads is not None
        return ALiBi(nheads, hp.get("maxpos", 10000))
    if name in ("T5RPE", "Rotery", "Rotary"):
        raise NotImplementedError(f"{name} is outside the VAE-GSLM hot path (no shipped config selects it)")
    raise ValueError(f"{name} is not a valid PE type.")
