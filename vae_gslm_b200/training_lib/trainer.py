""""Random init" of the reference (``training_lib/trainer.py:113-125``): PyTorch defaults, then every bias
to zero and the modules' own ``custom_weight_init`` (attention projections U(±1/√(dim/3)), embedding U(±1))."""
import torch
import torch.nn as nn


def init_weights(module: nn.Module, init_std: float = 1.0) -> None:
    """use as ``model.apply(init_weights)``"""
    bias = getattr(module, "bias", None)
    if isinstance(bias, torch.Tensor):
        with torch.no_grad():
            bias.zero_()
    if hasattr(module, "custom_weight_init"):
        module.custom_weight_init(init_std)
