"""(tensor, accumulated log-determinant) pair passed between coupling layers (reference ``modules/flow/utils.py``)."""
from typing import NamedTuple, Union

import torch

from ...utils.tensormask import TensorMask


class TensorLogdet(NamedTuple):
    tensor: Union[TensorMask, torch.Tensor]
    logdet: Union[float, torch.Tensor]
