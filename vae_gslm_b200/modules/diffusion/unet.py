"""Noise-prediction network of the diffusion decoder (reference ``modules/diffusion/unet.py:10-26,67-93``)."""
import torch
import torch.nn as nn

from ...hparams.hp import Hparams
from ...utils.tensormask import TensorMask
from ..activations import get_activation
from ..conv.layers import BottleNeckResNet
from ..position.absolute import SinCos


class TimeEmbedding(nn.Module):
    def __init__(self, hp: Hparams):
        super().__init__()
        hp.check_arg_in_hparams("activation", "maxpos", "dim")
        self.n_channels = hp.dim
        self.lin1 = nn.Linear(hp.dim, hp.dim, bias=hp.get("bias", True))
        self.act = get_activation(hp.activation)
        self.lin2 = nn.Linear(hp.dim, hp.dim, bias=hp.get("bias", True))
        self.embedding = SinCos(hp.dim, maxpos=hp.maxpos)

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        return self.lin2(self.act(self.lin1(self.embedding.get(t))))


class ConditionalBottleNeckUNet(nn.Module):
    def __init__(self, cond_dim: int, noise_dim: int, hp: Hparams):
        super().__init__()
        hp.check_arg_in_hparams("unet", "time_embedding")
        hp.unet.check_arg_in_hparams("conditional")
        hp.unet.time_dim = hp.time_embedding.dim
        self.cond_net = nn.Linear(cond_dim, hp.unet.condition_dim)
        self.time_embedding = TimeEmbedding(hp.time_embedding)
        self.unet = BottleNeckResNet(hp.unet, input_dim=noise_dim, output_dim=noise_dim)

    def forward(self, noise: TensorMask, t: torch.Tensor, cond: TensorMask) -> TensorMask:
        """noise, cond: [B,T,C]; t: [B] diffusion step indices."""
        from ... import ops
        temb = self.time_embedding(t)
        c = ops.linear(cond.value.to(self.unet.compute_dtype), self.cond_net.weight, self.cond_net.bias,
                       row_mask=cond.mask)                       # Linear + apply_mask in one GEMM epilogue
        return self.unet(noise, TensorMask(c, cond.mask), temb)


class ConditionalUNet(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("ConditionalUNet (ResNet body) is not selected by the VAE-GSLM configuration")
