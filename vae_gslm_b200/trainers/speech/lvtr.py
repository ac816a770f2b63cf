"""Loss assembly of the VAE-GSLM training step (reference ``trainers/speech/lvtr.py:103-145``) and a
minimal single-process training step around it.  The Lightning shell of the reference is out of scope
(SURVEY §2 row 15); data-parallel plumbing lives in ``vae_gslm_b200.dp``.
"""
from __future__ import annotations

from typing import Mapping, Optional

import torch

from ...training_lib.losses import masked_loss
from ...utils.tensormask import TensorMask


def kld_weight_at(global_step: int, kld_scale: float, warmup_kld: int = 0, zero_kld: int = 0) -> float:
    """KL weight schedule (:104-110): linear warm-up over `warmup_kld` steps after `zero_kld` zero steps.
    NB: with zero_kld = 0 and warmup_kld > 0 the weight is exactly 0 at global_step 0."""
    w = kld_scale
    if warmup_kld > 0 and zero_kld < (global_step + 1) <= warmup_kld:
        w = kld_scale * (global_step - zero_kld) / warmup_kld
    if zero_kld > 0 and global_step <= zero_kld:
        w = 0.0
    return w


def assemble_loss(output: Mapping, kld_weight: float, rec_loss_scale: float = 1.0, entropy_weight: float = 1.0,
                  token_kld_weight: float = 0.5, use_fused_kl: bool = True) -> Mapping[str, torch.Tensor]:
    """loss = rec·scale + kld·kw + ce·token_kld_weight·kw — every term a SUM over valid frames (:122-130)."""
    if use_fused_kl and entropy_weight == 1.0 and "kl_sum" in output:
        kld = output["kl_sum"]                       # computed inside the latent_back kernel
    else:
        kld = masked_loss(output["log_q"] * entropy_weight, output["log_p"], fn=lambda a, b: a - b)
    rec = output["decoder_output"]
    loss = rec * rec_loss_scale + kld * kld_weight
    ce = output.get("ce_loss")
    if ce is not None:
        loss = loss + ce * token_kld_weight * kld_weight
    return {"loss": loss, "kld": kld, "rec_loss": rec, "token_kld": ce, "kld_weight": kld_weight,
            "length": output["log_p"].mask.sum()}


def make_model_input(tokens: TensorMask, mel: TensorMask) -> TensorMask:
    """channel-interleaved model input [B,T,1+n_mels]: token id as float ⊕ mel (:117-118)."""
    return TensorMask(tokens.value.to(mel.value.dtype), tokens.mask).expand().cat(mel)
