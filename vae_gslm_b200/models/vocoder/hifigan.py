"""HiFi-GAN generator, inference only: mel frames → waveform, the last stage of `scripts.infer` after the diffusion decoder
(reference ``models/vocoder/hfgan.py:43-163`` generator, ``models/vocoder/vocoder.py:35-67`` wrapper; SURVEY §8f-4 tail).

Not on the measured hot path and not a hand-written kernel: the transposed / dilated convolutions run on cuDNN through
torch.  What is kept exactly is the reference's arithmetic (activation slopes — 0.1 inside, torch's default 0.01 before
``conv_post`` —, the per-scale average of the residual stacks, no masking inside the generator, lengths × Π upsample rates)
and its checkpoint format: a reference checkpoint is saved BEFORE ``remove_weight_norm`` and therefore holds the
weight-norm parametrisation (``…parametrizations.weight.original0/1`` or the older ``…weight_g/weight_v``); it is folded into
plain weights once, at load time, so that generation runs without the per-call ``g·v/‖v‖`` kernels.
"""
from __future__ import annotations

import os
from typing import Dict, Mapping, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...hparams.hp import Hparams
from ...utils.helpers import get_padding
from ...utils.tensormask import TensorMask

LRELU_SLOPE = 0.1


def fold_weight_norm(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """state dict with weight-norm parametrisations → plain ``.weight`` tensors (w = g · v / ‖v‖, norm over every
    dimension but the first — torch's ``weight_norm(dim=0)`` for Conv1d and ConvTranspose1d alike)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        for g_suffix, v_suffix in ((".parametrizations.weight.original0", ".parametrizations.weight.original1"),
                                   (".weight_g", ".weight_v")):
            if k.endswith(g_suffix):
                base = k[:-len(g_suffix)]
                direction = sd[base + v_suffix]
                norm = direction.flatten(1).norm(dim=1).view(-1, *([1] * (direction.dim() - 1)))
                out[base + ".weight"] = v * direction / norm
                break
            if k.endswith(v_suffix):
                break
        else:
            out[k] = v
    return out


class ResBlock(nn.Module):
    """three (dilated conv → conv) residual pairs (hfgan.py:43-80)."""

    def __init__(self, channels: int, kernel_size: int, dilation) -> None:
        super().__init__()
        self.convs1 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                               padding=get_padding(kernel_size, d)) for d in dilation])
        self.convs2 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=1,
                                               padding=get_padding(kernel_size, 1)) for _ in dilation])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for c1, c2 in zip(self.convs1, self.convs2):
            x = c2(F.leaky_relu(c1(F.leaky_relu(x, LRELU_SLOPE)), LRELU_SLOPE)) + x
        return x


class Generator(nn.Module):
    def __init__(self, hp: Hparams) -> None:
        super().__init__()
        hp.check_arg_in_hparams("resblock_kernel_sizes", "upsample_rates", "in_channels", "upsample_initial_channel",
                                "kernel_size", "upsample_kernel_sizes", "resblock_dilation_sizes")
        self.hp = hp
        self.num_kernels = len(hp.resblock_kernel_sizes)
        c0 = hp.upsample_initial_channel
        self.conv_pre = nn.Conv1d(hp.in_channels, c0, hp.kernel_size, 1, padding=get_padding(hp.kernel_size, 1))
        self.ups = nn.ModuleList([
            nn.ConvTranspose1d(c0 // 2 ** i, c0 // 2 ** (i + 1), k, u, padding=u // 2 + u % 2, output_padding=u % 2)
            for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes))])
        self.resblocks = nn.ModuleList([ResBlock(c0 // 2 ** (i + 1), k, d) for i in range(len(self.ups))
                                        for k, d in zip(hp.resblock_kernel_sizes, hp.resblock_dilation_sizes)])
        self.conv_post = nn.Conv1d(c0 // 2 ** len(self.ups), 1, hp.kernel_size, 1, padding=get_padding(hp.kernel_size, 1))
        self.total_upsample = 1
        for u in hp.upsample_rates:
            self.total_upsample *= u

    def load_reference_state_dict(self, sd: Mapping[str, torch.Tensor]) -> None:
        self.load_state_dict(fold_weight_norm(sd))

    @torch.no_grad()
    def forward(self, x: TensorMask) -> TensorMask:
        new_length = TensorMask.resize_length(x.length, self.total_upsample)
        h = self.conv_pre(x.value.transpose(-1, -2))
        for i, up in enumerate(self.ups):
            h = up(F.leaky_relu(h, LRELU_SLOPE))
            blocks = self.resblocks[i * self.num_kernels:(i + 1) * self.num_kernels]
            acc = blocks[0](h)
            for blk in blocks[1:]:
                acc = acc + blk(h)
            h = acc / self.num_kernels
        h = torch.tanh(self.conv_post(F.leaky_relu(h))).squeeze(1)        # default slope 0.01 here (hfgan.py:150)
        return TensorMask.fromlength(h, new_length)


class HiFiGAN(nn.Module):
    """``Vocoder`` of the reference for mel input: un-rescale, generate, mask (vocoder.py:35-67)."""

    def __init__(self, hp: Hparams, hp_rescale: Optional[Hparams] = None) -> None:
        super().__init__()
        self.hp = hp.feature
        self.hp_rescale = hp_rescale
        self.model = Generator(hp.model.generator)

    def match_spec(self, hp: Hparams) -> bool:
        return hp == self.hp

    @torch.no_grad()
    def decode(self, signal: TensorMask) -> TensorMask:
        if self.hp_rescale is not None:
            signal = TensorMask(signal.value * self.hp_rescale.std + self.hp_rescale.mean, signal.mask).apply_mask()
        return self.model(signal).apply_mask()

    @classmethod
    def from_pretrained(cls, path: str, **kwargs) -> "HiFiGAN":
        hp = Hparams.from_yamlfile(os.path.join(path, "hp.yaml"))
        hp.check_arg_in_hparams("model", "feature")
        hp.model.check_arg_in_hparams("generator")
        model = cls(hp, **kwargs)
        model.model.load_reference_state_dict(torch.load(os.path.join(path, "last-cpt.ckpt"), weights_only=True))
        return model.eval()
