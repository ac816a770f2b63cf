"""Convolutional blocks of the posterior encoder, the diffusion UNet and the utterance encoder
(reference ``modules/conv/layers.py``: Conv1d :13-31, ResidualBlock family :70-295,
BottleNeckResNet :386-540, ConvNormAct / CNNStack :543-652).

SURVEY §8f-1 (first "next" row) — the residual blocks now run on libvgslm:
  * activations stay ``[B,T,C]`` (the reference transposes to B,C,T for cuDNN and back);
  * ``norm(conv1(x) [+ time_emb])`` — depthwise k=7 conv (causal or future padded) + channel "InstanceNorm" — is ONE
    fused kernel (``ops.dwconv_ln``);
  * the 1x1 convolutions ``conv2`` / ``conv3`` / ``skip_conv`` are plain GEMMs on the same rows (``ops.linear`` with
    the Conv1d weight ``[N,K,1]`` viewed as ``[N,K]``), with bias + ReLU/SiLU and the residual add fused in the epilogue.
Parameter names/shapes are the reference's (checkpoints load unchanged).  Its quirk that activations in padded frames
are NOT zeroed between blocks is kept.  The small utterance encoder (strided convs over ≤200
frames) stays on torch/cuDNN.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...hparams.hp import Hparams
from ...utils.helpers import get_padding
from ...utils.tensormask import TensorMask
from ..activations import get_activation
from ..norm import InstanceNorm, get_norm_fn


class Conv1d(nn.Conv1d):
    """nn.Conv1d that also accepts an asymmetric (left, right) zero padding."""

    def __init__(self, *args, **kwargs):
        pad = kwargs.get("padding", 0)
        self.two_side_padding = None
        if isinstance(pad, tuple):
            assert len(pad) == 2
            self.two_side_padding = pad
            kwargs["padding"] = 0
        super().__init__(*args, **kwargs)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.two_side_padding is not None:
            x = F.pad(x, self.two_side_padding)
        return super().forward(x)


class ResidualBlock(nn.Module):
    """x + conv3(act(conv2([norm(conv1(x) + t) ; cond])))  on a B,T,C TensorMask."""

    def __init__(self, hp: Hparams):
        super().__init__()
        hp.check_arg_in_hparams("in_channels", "hidden_channels", "kernel_size", "norm", "activation")
        assert hp.norm.identifier != "LayerNorm", "BCT format not supported"
        if hp.get("shortcut", False) or hp.has("layer_scale") or hp.get("dropout", 0.0):
            raise NotImplementedError("shortcut / layer_scale / dropout are unused by the VAE-GSLM configuration")
        ch = hp.in_channels
        padding = get_padding(hp.kernel_size, causal=hp.get("causal_padding", False),
                              future=hp.get("future_padding", False))
        self.pad_left = padding[0] if isinstance(padding, tuple) else padding
        self.norm = get_norm_fn(ch, hp.norm)
        if not isinstance(self.norm, InstanceNorm):
            raise NotImplementedError("the fused conv block implements the channel 'InstanceNorm' of the configuration")
        self.act = get_activation(hp.activation)
        self._act_id = ops.ACT_IDS.get(hp.activation.identifier)
        if self._act_id is None:
            raise NotImplementedError(f"activation {hp.activation.identifier} is not fused (ReLU / GELU / SiLU are)")
        self.conv1 = Conv1d(ch, ch, kernel_size=hp.kernel_size, padding=padding, groups=ch)
        self.conv2 = nn.Conv1d(ch + hp.get("aux_in_channels", 0), hp.hidden_channels, kernel_size=1)
        self.conv3 = nn.Conv1d(hp.hidden_channels, ch, kernel_size=1)

    def _body(self, x: TensorMask, t_add: Optional[torch.Tensor] = None,
              cond: Optional[torch.Tensor] = None) -> TensorMask:
        xv = x.value
        h = ops.dwconv_ln(xv, self.conv1.weight, self.conv1.bias, t_add, self.norm.weight, self.norm.bias,
                          self.pad_left, self.norm.eps)
        if cond is not None:
            h = torch.cat([h, cond.to(h.dtype)], -1)
        a = ops.linear(h, self.conv2.weight, self.conv2.bias, act=self._act_id)
        y = ops.linear(a, self.conv3.weight, self.conv3.bias, residual=xv)
        return TensorMask(y, x.mask)

    def forward(self, x: TensorMask) -> TensorMask:
        return self._body(x)


def _setup_condition(block: ResidualBlock, hp: Hparams) -> None:
    block.condition_type = hp.get("condition_type", "film")
    if block.condition_type == "film":
        raise NotImplementedError("FiLM-conditioned conv blocks are unused by the VAE-GSLM configuration")
    hp.aux_in_channels = hp.get("in_dim", hp.in_channels)


class ConditionalResidualBlock(ResidualBlock):
    def __init__(self, hp: Hparams):
        _setup_condition(self, hp)
        super().__init__(hp)

    def forward(self, x: TensorMask, c: TensorMask) -> TensorMask:
        return self._body(x, cond=c.value)


class TemporalResidualBlock(ResidualBlock):
    def __init__(self, hp: Hparams):
        super().__init__(hp)
        hp.check_arg_in_hparams("time_dim")
        self.time_emb = nn.Linear(hp.time_dim, hp.in_channels)

    def forward(self, x: TensorMask, t: torch.Tensor) -> TensorMask:
        return self._body(x, t_add=self.time_emb(self.act(t.float())))


class TCResidualBlock(ResidualBlock):
    def __init__(self, hp: Hparams):
        _setup_condition(self, hp)
        super().__init__(hp)
        hp.check_arg_in_hparams("time_dim")
        self.time_emb = nn.Linear(hp.time_dim, hp.in_channels)

    def forward(self, x: TensorMask, c: TensorMask, t: torch.Tensor) -> TensorMask:
        return self._body(x, t_add=self.time_emb(self.act(t.float())), cond=c.value)


class BottleNeckResNet(nn.Module):
    """Linear in → N residual blocks (optionally time/condition aware, with UNet-style skips) →
    channel norm → Linear out.  Input/output are B,T,C TensorMasks; nothing is transposed."""

    def __init__(self, hp: Hparams, input_dim: Optional[int] = None, output_dim: Optional[int] = None) -> None:
        super().__init__()
        self.hp = hp
        hp.check_arg_in_hparams("num_layers", "layer", "init_channel", "out_channels", "hidden_channels",
                                "resample_rates", "resample_ksize")
        n = hp.num_layers
        boundary = hp.upward_layer.boundary if hp.has("upward_layer") else n
        assert boundary <= n
        in_channels = ([hp.init_channel] + list(hp.out_channels))[:-1]
        if hp.has("conditional"):
            hp.check_arg_in_hparams("condition_dim")
            hp.layer.in_dim = hp.condition_dim
            if hp.has("upward_layer"):
                hp.upward_layer.in_dim = hp.condition_dim
        self.conditional = hp.get("conditional", [False] * n)
        self.time_dim = hp.get("time_dim", None)
        self.skip_connection = hp.get("skip_connection", [None] * n)
        self.skip_concat = hp.get("connection_type", None) == "concat"
        for seq in (hp.resample_rates, hp.resample_ksize, hp.out_channels, hp.hidden_channels, self.skip_connection):
            assert len(seq) == n
        layers, samples, skip_conv = [], [], []
        for i in range(n):
            lhp = hp.layer if i < boundary else hp.upward_layer
            lhp.in_channels, lhp.hidden_channels, lhp.aux_in_channels = in_channels[i], hp.hidden_channels[i], 0
            skip_conv.append(nn.Conv1d(2 * in_channels[i], in_channels[i], 1)
                             if self.skip_connection[i] is not None and self.skip_concat else nn.Identity())
            timed = self.time_dim is not None
            if timed:
                lhp.time_dim = self.time_dim
            if self.conditional[i]:
                layers.append(TCResidualBlock(lhp) if timed else ConditionalResidualBlock(lhp))
            else:
                layers.append(TemporalResidualBlock(lhp) if timed else ResidualBlock(lhp))
            if hp.resample_rates[i] not in (1, -1):
                raise NotImplementedError("resampling blocks are unused by the VAE-GSLM configuration")
            assert in_channels[i] == hp.out_channels[i]
            samples.append(nn.Identity())
        self.layers = nn.ModuleList(layers)
        self.samples = nn.ModuleList(samples)
        self.skip_conv = nn.ModuleList(skip_conv)
        self.linear = nn.Linear(input_dim, hp.init_channel) if input_dim is not None else None
        self.out_linear = nn.Linear(hp.out_channels[-1], output_dim) if output_dim is not None else None
        self.final_norm = get_norm_fn(hp.out_channels[-1], hp.layer.norm) if hp.get("final_norm", False) else None
        if hp.get("first_norm", False):
            raise NotImplementedError("first_norm is unused by the VAE-GSLM configuration")
        self.first_norm = None
        self.compute_dtype = torch.float32          # set by LVTR.set_compute_dtype

    def forward(self, x: TensorMask, c: Optional[TensorMask] = None, t: Optional[torch.Tensor] = None) -> TensorMask:
        mask = x.mask
        h = x.value.to(self.compute_dtype)
        if self.linear is not None:
            h = ops.linear(h, self.linear.weight, self.linear.bias, row_mask=mask)
        cur = TensorMask(h, mask)
        cond = TensorMask(c.value.to(self.compute_dtype), c.mask) if c is not None else None
        records = [cur]
        for layer, is_cond, skip, merge in zip(self.layers, self.conditional, self.skip_connection, self.skip_conv):
            args = ([cond] if is_cond else []) + ([t] if self.time_dim is not None else [])
            cur = layer(cur, *args)
            if skip is not None:
                if self.skip_concat:
                    both = torch.cat([cur.value, records[skip].value], -1)
                    cur = TensorMask(ops.linear(both, merge.weight, merge.bias), mask)
                else:
                    cur = cur + records[skip]
            records.append(cur)
        h = cur.value
        if self.final_norm is not None:
            h = ops.dwconv_ln(h, None, None, None, self.final_norm.weight, self.final_norm.bias, 0,
                              self.final_norm.eps)
        if self.out_linear is not None and self.out_linear.out_features < 8:
            # 512 → 4 projection of the posterior encoder: too narrow for the GEMM tiles, a 16 KB weight — torch
            h = torch.where(mask[..., None], F.linear(h.float(), self.out_linear.weight, self.out_linear.bias), 0.0)
        elif self.out_linear is not None:
            h = ops.linear(h, self.out_linear.weight, self.out_linear.bias, row_mask=mask)
        else:
            h = ops.mask_rows(h, mask)
        return TensorMask(h, mask)

    @property
    def sample_ratio(self) -> float:
        ratio = 1.0
        for r in self.hp.resample_rates:
            ratio = ratio * r if r > 0 else ratio / -r
        return ratio


class ConvNormAct(nn.Module):
    """(strided) conv → channel norm → activation (torch/cuDNN, B,C,T).

    Reference quirk kept on purpose (layers.py:576,591-593): ``self.stride`` holds 1/stride and the valid
    length is resized by 1/self.stride = stride, so after a stride-2 layer the recorded length DOUBLES
    while T halves — in practice every down-sampled frame of the utterance encoder counts as valid."""

    def __init__(self, hp: Hparams):
        super().__init__()
        hp.check_arg_in_hparams("in_channels", "out_channels", "kernel_size", "stride", "norm", "activation")
        assert hp.norm.identifier != "LayerNorm", "BCT format not supported"
        if hp.stride > 1:
            raise NotImplementedError("transposed (up-sampling) ConvNormAct is unused by the VAE-GSLM configuration")
        stride = -hp.stride if hp.stride < 0 else hp.stride
        padding = get_padding(hp.kernel_size, causal=hp.get("causal_padding", False),
                              future=hp.get("future_padding", False))
        self.norm = get_norm_fn(hp.out_channels, hp.norm)
        self.act = get_activation(hp.activation)
        self.conv = Conv1d(hp.in_channels, hp.out_channels, kernel_size=hp.kernel_size, stride=stride, padding=padding)
        self.stride = 1.0 / float(stride)

    def forward(self, x: TensorMask) -> TensorMask:
        h = self.act(self.norm(self.conv(x.value)))
        if self.stride != 1:
            return TensorMask.fromlength(h, TensorMask.resize_length(x.length, 1.0 / float(self.stride)), axis=2)
        return TensorMask(h, x.mask, axis=2)


class CNNStack(nn.Module):
    def __init__(self, hp: Hparams, input_dim: Optional[int] = None, output_dim: Optional[int] = None) -> None:
        super().__init__()
        self.hp = hp
        hp.check_arg_in_hparams("num_layers", "layer", "init_channel", "out_channels", "resample_rates",
                                "resample_ksize")
        in_channels = ([hp.init_channel] + list(hp.out_channels))[:-1]
        for seq in (hp.resample_rates, hp.resample_ksize, hp.out_channels):
            assert len(seq) == hp.num_layers
        layers = []
        for i in range(hp.num_layers):
            lhp = hp.layer
            lhp.in_channels, lhp.out_channels = in_channels[i], hp.out_channels[i]
            lhp.kernel_size, lhp.stride = hp.resample_ksize[i], hp.resample_rates[i]
            layers.append(ConvNormAct(lhp))
        self.layers = nn.ModuleList(layers)
        self.linear = nn.Linear(input_dim, hp.init_channel) if input_dim is not None else None
        self.out_linear = nn.Linear(hp.out_channels[-1], output_dim) if output_dim is not None else None
        self.compute_dtype = torch.float32          # set by LVTR.set_compute_dtype

    def _rows_path(self, x: TensorMask) -> bool:
        """the libvgslm route: strided Conv1d → channel norm → ReLU layers on a CUDA [B,T,C] tensor"""
        from ..norm import InstanceNorm
        return (x.value.is_cuda and self.linear is not None and self.out_linear is not None
                and all(isinstance(l.norm, InstanceNorm) and isinstance(l.act, nn.ReLU) and l.stride != 1
                        and l.conv.two_side_padding is None and l.conv.dilation[0] == 1 and l.conv.groups == 1
                        and l.conv.out_channels >= 16 and l.conv.out_channels % 8 == 0       # the row-norm kernel's range
                        and (l.conv.in_channels * l.conv.kernel_size[0]) % 8 == 0
                        for l in self.layers))

    def forward(self, x: TensorMask) -> TensorMask:
        if self._rows_path(x):
            return self._forward_rows(x)
        if self.linear is not None:
            x = TensorMask(self.linear(x.value), x.mask).apply_mask()
        x = x.transpose()
        for layer in self.layers:
            x = layer(x)
        x = x.transpose()
        if self.out_linear is not None:
            x = TensorMask(self.out_linear(x.value), x.mask).apply_mask()
        return x.apply_mask()

    def _forward_rows(self, x: TensorMask) -> TensorMask:
        """Same arithmetic as above without leaving the [B,T,C] layout (reference: conv/layers.py:549-593,631-642): every
        strided convolution is a window gather (ops.im2col, which also applies the previous layer's ReLU on read) plus ONE
        tcgen05 GEMM with the Conv1d weight used in place, the channel norm is the fused row kernel of the encoder
        (ops.dwconv_ln without its convolution), and nothing runs on cuDNN.  The recorded length GROWS by the stride per
        layer exactly as in the reference (ConvNormAct quirk), so every down-sampled frame is valid."""
        h = ops.linear(x.value.to(self.compute_dtype), self.linear.weight, self.linear.bias, row_mask=x.mask)
        length = x.length
        relu = False
        for layer in self.layers:
            conv = layer.conv
            a = ops.im2col(h, conv.kernel_size[0], conv.stride[0], conv.padding[0], relu=relu)
            y = ops.linear(a, conv.weight, conv.bias)
            h = ops.dwconv_ln(y, None, None, None, layer.norm.weight, layer.norm.bias, 0, layer.norm.eps)
            relu = True
            length = TensorMask.resize_length(length, 1.0 / float(layer.stride))
        h = ops.im2col(h, 1, 1, 0, relu=True)
        out = TensorMask.fromlength(h, length, axis=1)
        y = ops.linear(h, self.out_linear.weight, self.out_linear.bias, row_mask=out.mask)
        return TensorMask(y, out.mask)

    @property
    def sample_ratio(self) -> float:
        ratio = 1.0
        for r in self.hp.resample_rates:
            ratio = ratio * r if r > 0 else ratio / -r
        return ratio
