"""Activation factory (reference ``modules/activations.py:5-18``).  The fused kernels know ReLU and exact
GELU (``ops.ACT_IDS``); anything else runs as a plain torch module outside the fused epilogues."""
import torch.nn as nn

from ..hparams.hp import Hparams

_TABLE = {
    "ReLU": lambda hp: nn.ReLU(),
    "SELU": lambda hp: nn.SELU(),
    "GELU": lambda hp: nn.GELU(),
    "LeakyRELU": lambda hp: nn.LeakyReLU(negative_slope=hp.slope),
    "SiLU": lambda hp: nn.SiLU(),
}


def get_activation(hp: Hparams) -> nn.Module:
    try:
        return _TABLE[hp.identifier](hp)
    except KeyError:
        raise ValueError(f"{hp.identifier} not in the usable activation function lists.") from None
