"""ORACLE — CPU/any-device restatement of the reference's VAE-GSLM hot path in plain PyTorch fp32.

THIS FILE IS TEST INFRASTRUCTURE.  It is the checker for the CUDA path, never the thing shipped or
measured: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import it.  Nothing under ``vae_gslm_b200/`` imports it.

It is a flat, functional restatement (state_dict in, tensors out) of the reference modules, each block
citing the reference lines it follows.  All arithmetic is torch fp32 (``F.linear``, explicit softmax
attention with the dense additive mask the reference builds, ``F.conv1d``); there are no custom kernels.

Parity pin: the reference ships no tests or golden vectors, so the pin is differential —
``tests/golden/make_golden.py`` imports the REAL reference from /root/reference in the build container,
runs it with injected RNG tensors on a small configuration, and commits inputs + outputs + gradients
as fixtures; ``tests/test_oracle_golden.py`` checks this file against those fixtures to ~1e-6.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)

# The bf16 gradient bound asserted by tests/test_model_gpu.py, tests/test_parity_full_gpu.py and __graft_entry__.smoke()
# (Frobenius-relative error of a parameter's gradient against this oracle's fp32 gradient): within the north star's 2e-2,
# or no noisier than BF16_VS_AUTOCAST x what torch.autocast(bf16) does to the same gradient of this oracle on the same GPU.
# profiles/r02_parity_bf16.md lists all three numbers for every tensor of the full configuration (there the ratio is <= 1.05 at
# the BASELINE shapes; the factor 2.5 covers the small test shapes — B=2, T=160 and the golden configuration — where the bf16 error of
# single tensors of a few hundred elements fluctuates by 2x between two bf16 evaluations of the same gradient).
BF16_GRAD_BOUND = 2e-2
BF16_VS_AUTOCAST = 2.5


def bf16_grad_within_bound(err: float, err_autocast: float) -> bool:
    return err <= max(BF16_GRAD_BOUND, BF16_VS_AUTOCAST * err_autocast)


# ----------------------------------------------------------------------------------------- helpers
def _mask3(v: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """TensorMask.apply_mask (utils/tensormask.py:63-67) for [B,T,C]."""
    return torch.where(mask[..., None], v, torch.zeros((), dtype=v.dtype, device=v.device))


def _lin(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def alibi_slopes(n: int) -> List[float]:
    """position/alibi.py:19-30"""
    def pow2(n):
        start = 2 ** (-2 ** -(math.log2(n) - 3))
        return [start * start ** i for i in range(n)]
    if math.log2(n).is_integer():
        return pow2(n)
    c = 2 ** math.floor(math.log2(n))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: n - c]


def rmsnorm(x: torch.Tensor, scale: torch.Tensor, eps: float) -> torch.Tensor:
    """modules/norm.py:28-32"""
    x = x.float()
    return scale * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def channel_ln_bct(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    """'InstanceNorm' of modules/norm.py:43-47: per-(b,t) LayerNorm over channels of a B,C,T tensor, UNBIASED variance."""
    x = x.float()
    var, mean = torch.var_mean(x, dim=1, keepdim=True)
    return w[:, None] * ((x - mean) * torch.rsqrt(var + eps)) + b[:, None]


_ACT = {"ReLU": F.relu, "GELU": F.gelu, "SiLU": F.silu, "SELU": F.selu}


# ------------------------------------------------------------------------------- conv networks
def _res_block(sd, p, x_bct, act, eps, pad, t_add=None, cond_bct=None):
    """ResidualBlock family, modules/conv/layers.py:117-135,181-193,238-253,275-295"""
    w1 = sd[p + ".conv1.weight"]
    h = F.conv1d(F.pad(x_bct, pad), w1, sd[p + ".conv1.bias"], groups=w1.shape[0])
    if t_add is not None:
        h = h + t_add
    h = channel_ln_bct(h, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps)
    if cond_bct is not None:
        h = torch.cat([h, cond_bct], 1)
    h = act(F.conv1d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"]))
    h = F.conv1d(h, sd[p + ".conv3.weight"], sd[p + ".conv3.bias"])
    return h + x_bct


def bottleneck_resnet(sd, p, hp, x, mask, cond=None, temb=None):
    """BottleNeckResNet.forward, modules/conv/layers.py:490-530 (rates all 1)."""
    n = hp["num_layers"]
    boundary = hp["upward_layer"]["boundary"] if "upward_layer" in hp else n
    conditional = hp.get("conditional", [False] * n)
    skips = hp.get("skip_connection", [None] * n)
    h = _mask3(_lin(sd, p + ".linear", x), mask).transpose(1, 2)          # B,C,T
    cond_bct = cond.transpose(1, 2) if cond is not None else None
    records = [h]
    for i in range(n):
        lhp = hp["layer"] if i < boundary else hp["upward_layer"]
        k = lhp["kernel_size"]
        pad = (k - 1, 0) if lhp.get("causal_padding", False) else ((0, k - 1) if lhp.get("future_padding", False)
                                                                   else ((k - 1) // 2, (k - 1) // 2))
        act = _ACT[lhp["activation"]["identifier"]]
        lp = f"{p}.layers.{i}"
        t_add = None
        if temb is not None:
            t_add = _lin(sd, lp + ".time_emb", act(temb))[..., None]
        h = _res_block(sd, lp, h, act, lhp["norm"]["eps"], pad, t_add, cond_bct if conditional[i] else None)
        if skips[i] is not None:
            h = F.conv1d(torch.cat([h, records[skips[i]]], 1), sd[f"{p}.skip_conv.{i}.weight"],
                         sd[f"{p}.skip_conv.{i}.bias"])
        records.append(h)
    if hp.get("final_norm", False):
        h = channel_ln_bct(h, sd[p + ".final_norm.weight"], sd[p + ".final_norm.bias"], hp["layer"]["norm"]["eps"])
    h = h.transpose(1, 2)
    return _mask3(_mask3(_lin(sd, p + ".out_linear", h), mask), mask)


def utterance_encoder(sd, p, hp, x, mask):
    """CNNStack + TimeAggregation, modules/conv/layers.py:631-642,600-613; linear/layers.py:260-262"""
    h = _mask3(_lin(sd, p + ".0.linear", x), mask).transpose(1, 2)
    length = mask.long().sum(-1)
    act = _ACT[hp["layer"]["activation"]["identifier"]]
    for i, (k, rate) in enumerate(zip(hp["resample_ksize"], hp["resample_rates"])):
        stride = -rate if rate < 0 else rate
        pad = int(((k - 1) + 1 - 1) / 2)                    # get_padding ignores the stride here (layers.py:557-558)
        h = F.conv1d(h, sd[f"{p}.0.layers.{i}.conv.weight"], sd[f"{p}.0.layers.{i}.conv.bias"], stride=stride,
                     padding=pad)
        h = act(channel_ln_bct(h, sd[f"{p}.0.layers.{i}.norm.weight"], sd[f"{p}.0.layers.{i}.norm.bias"],
                               hp["layer"]["norm"]["eps"]))
        if stride != 1:
            # reference quirk (layers.py:576,591-593): ConvNormAct stores self.stride = 1/stride and then resizes
            # the length by 1/self.stride = stride, i.e. the valid length GROWS while T shrinks
            length = torch.ceil(length.float() * float(stride)).long()
    m = torch.arange(h.shape[2], device=h.device)[None, :] < length[:, None]
    h = h.transpose(1, 2)
    h = _mask3(_mask3(_lin(sd, p + ".0.out_linear", h), m), m)
    return h.sum(1) / m.long().sum(-1)[:, None]


# --------------------------------------------------------------------------------- transformer
def self_attention(sd, p, x, kv_mask, nheads, slopes, past=None):
    """SelfAttention.forward, modules/attention/attention.py:52-85: dense key-padding ∧ causal mask as 0/−inf,
    plus ALiBi −slope·|i−j|, softmax(QKᵀ/√d + mask)V, out_proj, masked."""
    B, Tq, C = x.shape
    q, k, v = _lin(sd, p + ".in_proj", x).chunk(3, -1)
    if past is not None:
        k = torch.cat([past["key"], k], 1)
        v = torch.cat([past["value"], v], 1)
        kv_valid = torch.ones(B, k.shape[1], dtype=torch.bool, device=x.device)
    else:
        kv_valid = kv_mask
    Tk = k.shape[1]
    allow = kv_valid[:, None, :].expand(-1, Tk, -1) & torch.ones(Tk, Tk, dtype=torch.bool, device=x.device).tril()
    bias = torch.zeros(B, Tk, Tk, dtype=v.dtype, device=x.device).masked_fill_(~allow, float("-inf"))
    bias = bias[:, None].expand(-1, nheads, -1, -1)
    pos = torch.arange(Tk, device=x.device)
    rel = (pos[None, :] - pos[:, None]).abs()
    bias = bias + (-slopes.view(1, nheads, 1, 1) * rel[None, None])
    bias = bias[:, :, -Tq:]
    D = C // nheads
    qh, kh, vh = (t.view(B, t.shape[1], nheads, D).transpose(1, 2) for t in (q, k, v))
    w = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(D) + bias, -1)
    o = (w @ vh).transpose(1, 2).reshape(B, Tq, C)
    return _lin(sd, p + ".out_proj", o), {"key": k.detach(), "value": v.detach()}


def transformer_stack(sd, p, hp, x, mask, past_kv=None):
    """TransformerLayerStack.run + TransformerLayer.forward (pre-LN), modules/transformer/layers.py:41-93,134-195"""
    L = hp["num_layers"]
    lh = hp["layer"]
    nheads = lh["self_attn"]["nheads"]
    eps = lh["norm"]["eps"]
    act = _ACT[lh["activation"]["identifier"]]
    slopes = torch.tensor(alibi_slopes(nheads), dtype=torch.float32, device=x.device)
    h = _mask3(_lin(sd, p + ".linear", x), mask)
    kvs = []
    for i in range(L):
        lp = f"{p}.layers.{i}"
        n = _mask3(rmsnorm(h, sd[lp + ".norm1.scale"], eps), mask)
        a, kv = self_attention(sd, lp + ".self_attn", n, mask, nheads, slopes, None if past_kv is None else past_kv[i])
        kvs.append(kv)
        h = h + _mask3(a, mask)
        n = rmsnorm(h, sd[lp + ".norm3.scale"], eps)
        h = _mask3(h + _lin(sd, lp + ".linear2", act(_lin(sd, lp + ".linear1", n))), mask)
    return rmsnorm(h, sd[p + ".final_norm.scale"], eps), kvs


# ---------------------------------------------------------------------------------------- flow
def _coupling_stats(sd, p, hp, x0, c):
    """LinearCoupling conditioner, modules/flow/layers.py:49-65 (+ FiLM linear/layers.py:278-288)"""
    s = F.layer_norm(_lin(sd, p + ".linear1", x0), (sd[p + ".norm.weight"].shape[0],), sd[p + ".norm.weight"],
                     sd[p + ".norm.bias"], hp["layer"]["norm"]["eps"])
    gamma, beta = _lin(sd, p + ".film.linear", c).chunk(2, -1)
    s = gamma * s + beta
    m, logs = _lin(sd, p + ".linear2", _ACT[hp["layer"]["activation"]["identifier"]](s)).chunk(2, -1)
    _max, _min = hp["layer"]["scale_range"]                    # [0.5, 2.0] unpacked as (_max, _min): :62-65
    logs = torch.log(torch.sigmoid(logs) * (_max - _min) + _min)
    return m, logs


def flow_forward(sd, p, hp, z, mask, c):
    """CouplingStack.forward with every layer flip=True, modules/flow/layers.py:42-73,219,225-234"""
    x = z
    logdet = 0.0
    for i in range(hp["num_layers"]):
        x0, x1 = x.chunk(2, -1)
        x0, x1 = x1, x0
        m, logs = _coupling_stats(sd, f"{p}.layers.{i}", hp, x0, c)
        x = torch.cat([x0, m + x1 * torch.exp(logs)], -1)
        logdet = logdet + _mask3(logs, mask)
    return x, logdet


def flow_reverse(sd, p, hp, x, c):
    """CouplingStack.reverse, modules/flow/layers.py:75-99,236-245"""
    for i in reversed(range(hp["num_layers"])):
        x0, x1 = x.chunk(2, -1)
        m, logs = _coupling_stats(sd, f"{p}.layers.{i}", hp, x0, c)
        x1 = (x1 - m) * torch.exp(-logs)
        x = torch.cat([x1, x0], -1)
    return x


# ----------------------------------------------------------------------------------- diffusion
def sincos_table(maxpos: int, ndim: int, device) -> torch.Tensor:
    """position/absolute.py:13-21"""
    p = torch.zeros(maxpos, ndim)
    pi = torch.arange(maxpos).float().unsqueeze(1) * torch.exp(
        torch.arange(0, ndim, 2).float() * -(math.log(10000.0) / ndim))
    p[:, 0::2] = torch.sin(pi)
    p[:, 1::2] = torch.cos(pi)
    return p.to(device)


def diffusion_loss(sd, p, hp, mel_scaled, mask, cond, t, noise):
    """GaussianDiffusion1D.forward/p_losses/q_sample (ddpm.py:328-374) with ConditionalBottleNeckUNet (unet.py:67-93)
    and the masked L1 of losses.py:9-27,44-57."""
    sa = sd[p + ".sqrt_alphas_cumprod"].gather(-1, t).reshape(-1, 1, 1)
    s1 = sd[p + ".sqrt_one_minus_alphas_cumprod"].gather(-1, t).reshape(-1, 1, 1)
    x_t = _mask3(sa * mel_scaled + s1 * noise, mask)
    te = hp["cond_unet"]["time_embedding"]
    act = _ACT[te["activation"]["identifier"]]
    emb = sincos_table(te["maxpos"], te["dim"], mel_scaled.device)[t]
    temb = _lin(sd, p + ".model.time_embedding.lin2", act(_lin(sd, p + ".model.time_embedding.lin1", emb)))
    c = _mask3(_lin(sd, p + ".model.cond_net", cond), mask)
    pred = bottleneck_resnet(sd, p + ".model.unet", hp["cond_unet"]["unet"], x_t, mask, cond=c, temb=temb)
    target = _mask3(noise, mask)
    return (_mask3(pred, mask) - target).abs().mean(-1).sum(-1).sum()


def unet_predict(sd, p, hp, x_t, t, cond, mask):
    """ConditionalBottleNeckUNet.forward (unet.py:67-93): noise prediction for x_t at integer steps t [B]."""
    te = hp["cond_unet"]["time_embedding"]
    act = _ACT[te["activation"]["identifier"]]
    emb = sincos_table(te["maxpos"], te["dim"], x_t.device)[t]
    temb = _lin(sd, p + ".model.time_embedding.lin2", act(_lin(sd, p + ".model.time_embedding.lin1", emb)))
    c = _mask3(_lin(sd, p + ".model.cond_net", cond), mask)
    return _mask3(bottleneck_resnet(sd, p + ".model.unet", hp["cond_unet"]["unet"], x_t, mask, cond=c, temb=temb), mask)


def cosine_alphas_cumprod(timesteps: int, s: float = 0.008) -> torch.Tensor:
    """cosine_beta_schedule + cumprod (ddpm.py:127-138,171-180), float64 → float32 like the reference buffers."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    return torch.cumprod(1.0 - betas, dim=0).float()


def lvtr_decode(sd, hp, frames, mask, u_c, start, step_noise, sampling_steps):
    """LVTR.decode (lvtr.py:288-306) → GaussianDiffusion1D.ddim_sample (ddpm.py:284-321), pred_noise objective.
    ``start`` [B,T,n_mels] is the initial noise (torch.randn in the reference), ``step_noise`` the per-step
    torch.randn_like draws (one per step except the last)."""
    dh = hp["decoder"]["diffusion"]
    assert dh.get("objective", "pred_noise") == "pred_noise" and dh["beta_schedule"]["identifier"] == "cosine"
    T_total = dh["timesteps"]
    eta, lo_hi, sig = dh.get("ddim_sampling_eta", 0.0), dh.get("clamp_range", [-1.0, 1.0]), dh.get("sigma", 1.0)
    ac = cosine_alphas_cumprod(T_total).to(frames.device)
    ids = frames[..., 0].long()
    E = _mask3(F.embedding(ids, sd["token_embedding.weight"]), mask)
    fused = E + F.relu(_lin(sd, "token_fuser.linear", frames[..., 1:]))                 # fuse_inputs :390-392
    cond = _mask3(torch.cat([fused, u_c[:, None].expand(-1, fused.shape[1], -1)], -1), mask)
    times = list(reversed(torch.linspace(-1, T_total - 1, steps=sampling_steps + 1).int().tolist()))
    img = _mask3(start, mask)
    noises = list(step_noise)
    for t, t_next in zip(times[:-1], times[1:]):
        tc = torch.full((frames.shape[0],), t, dtype=torch.long, device=frames.device)
        eps_hat = unet_predict(sd, "decoder", hp["decoder"], img, tc, cond, mask)
        x0 = (1.0 / ac[t]).sqrt() * img - (1.0 / ac[t] - 1).sqrt() * eps_hat      # predict_start_from_noise
        x0 = _mask3(_mask3(x0, mask).clamp(lo_hi[0], lo_hi[1]), mask)
        if t_next < 0:
            img = x0
            continue
        a, a_next = ac[t], ac[t_next]
        sigma = eta * ((1 - a / a_next) * (1 - a_next) / (1 - a)).sqrt()
        c = (1 - a_next - sigma ** 2).sqrt()
        img = _mask3(x0 * a_next.sqrt() + c * eps_hat + sigma * noises.pop(0) * sig, mask)
    return img * dh.get("input_scale", 1.0)


def lvtr_likelihood(sd, hp, x, mask, init_state):
    """LVTR.likelihood (lvtr.py:337-388) with the temperature-0 posterior: per-utterance mean token log-probability."""
    ids = x[..., 0].long()
    mel = x[..., 1:]
    E = _mask3(F.embedding(ids, sd["token_embedding.weight"]), mask)
    q = _lin(sd, "encoder.1.mean", bottleneck_resnet(sd, "encoder.0", hp["encoder"], mel, mask))   # sample = mean
    fused = E + F.relu(_lin(sd, "token_fuser.linear", q))
    shifted = _mask3(torch.cat([init_state, fused], 1)[:, :-1], mask)
    H, _ = transformer_stack(sd, "transformer.0", hp["transformer"], shifted, mask)
    logits = _lin(sd, "token_predictor.linear", F.relu(_lin(sd, "token_spliter.linear", H)))
    lp = torch.log_softmax(logits, -1).gather(-1, ids[..., None])[..., 0]
    return torch.where(mask, lp, torch.zeros_like(lp)).sum(-1) / mask.sum(-1)


# ------------------------------------------------------------------------------------- the model
def lvtr_forward(sd: Dict[str, torch.Tensor], hp: dict, x: torch.Tensor, mask: torch.Tensor,
                 utterance: torch.Tensor, utt_mask: torch.Tensor, rng: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """LVTR.forward, models/speech/lvtr.py:143-225.  ``rng`` holds the five draws of the reference forward in its
    order: eps_q [B,T,L], init_state [B,1,E], eps_p (unused by the loss), diff_t [B], diff_noise [B,T,n_mels]."""
    Ld = hp["latent_dim"]
    ids = x[..., 0].long()                                                     # :151-153
    mel = x[..., 1:]
    E = _mask3(F.embedding(ids, sd["token_embedding.weight"]), mask)           # :154 ; linear/layers.py:150-152
    h_enc = bottleneck_resnet(sd, "encoder.0", hp["encoder"], mel, mask)       # :156
    mean_q = _lin(sd, "encoder.1.mean", h_enc)                                 # linear/layers.py:91,106
    logstd_q = _lin(sd, "encoder.1.logstd", h_enc)
    sample_q = mean_q + rng["eps_q"] * torch.exp(logstd_q.float())             # :114-128 (temperature 1)
    log_q = -logstd_q - 0.5 - HALF_LOG_2PI                                     # :158
    sample_q = _mask3(sample_q, mask)                                          # :160
    fused = E + F.relu(_lin(sd, "token_fuser.linear", sample_q))               # :166,390-392
    shifted = torch.cat([rng["init_state"], fused], 1)[:, :-1]                 # :167-168 push + pop(1)
    shifted = _mask3(shifted, mask)
    H, _ = transformer_stack(sd, "transformer.0", hp["transformer"], shifted, mask)   # :170
    c = F.relu(_lin(sd, "q_spliter.linear", H))                                # :171
    mean_p = _lin(sd, "transformer.1.mean", c)                                 # :172
    logstd_p = _lin(sd, "transformer.1.logstd", c)
    y, logdet = flow_forward(sd, "transformer_flow", hp["transformer"]["flow"], sample_q, mask, c)   # :182-184
    log_p = (logdet.sum(-1) / Ld)[..., None] - logstd_p - HALF_LOG_2PI         # :185-188
    log_p = log_p + -0.5 * (torch.exp(-2 * logstd_p) * (y - mean_p) ** 2)      # :189-190
    logits = _lin(sd, "token_predictor.linear", F.relu(_lin(sd, "token_spliter.linear", H)))       # :194-195
    tgt = torch.where(mask, ids, torch.full_like(ids, -100))
    ce = F.cross_entropy(_mask3(logits, mask).reshape(-1, logits.shape[-1]), tgt.reshape(-1), reduction="sum",
                         ignore_index=-100)                                    # :196 ; losses.py:30-41
    u_c = utterance_encoder(sd, "utterance_encoder", hp["utterance_encoder"], utterance, utt_mask)   # :205
    cond = torch.cat([fused, u_c[:, None].expand(-1, fused.shape[1], -1)], -1)  # :202-207
    scale = hp["decoder"]["diffusion"].get("input_scale", 1.0)
    rec = diffusion_loss(sd, "decoder", hp["decoder"], mel / scale, mask, cond, rng["diff_t"], rng["diff_noise"])
    log_p_m, log_q_m = _mask3(log_p, mask), _mask3(log_q, mask)
    kld = (log_q_m - log_p_m).mean(-1).sum(-1).sum()                           # trainers/speech/lvtr.py:122-124
    return {"log_p": log_p_m, "log_q": log_q_m, "decoder_output": rec, "sample_q": sample_q,
            "transformer_latent": H, "ce_loss": ce, "kld": kld, "logits": logits, "mean_q": mean_q,
            "logstd_q": logstd_q, "u_c": u_c, "flow_y": y}


def total_loss(out: Dict[str, torch.Tensor], kw: float, rec_scale: float = 1.0, token_w: float = 0.5) -> torch.Tensor:
    """trainers/speech/lvtr.py:125-130"""
    return out["decoder_output"] * rec_scale + out["kld"] * kw + out["ce_loss"] * token_w * kw


def lvtr_step(sd, hp, x, past_kv, eps, temperature=1.0, push_init_state=False, init_state=None):
    """LVTR.step, models/speech/lvtr.py:227-286, with the prior eps injected and the token chosen greedily
    (argmax of the logits the reference feeds to softmax → multinomial).  Returns output [B,t,1+L], kv, logits, latent."""
    ids = x[..., 0].long()
    z = x[..., 1:]
    u = F.embedding(ids, sd["token_embedding.weight"]) + F.relu(_lin(sd, "token_fuser.linear", z))   # :244-247
    if push_init_state:
        u = torch.cat([init_state, u], 1)                                      # :248-251
    mask = torch.ones(u.shape[:2], dtype=torch.bool, device=u.device)
    H, kvs = transformer_stack(sd, "transformer.0", hp["transformer"], u, mask, past_kv)   # :253-257
    c = F.relu(_lin(sd, "q_spliter.linear", H))                                # :267
    mean_p, logstd_p = _lin(sd, "transformer.1.mean", c), _lin(sd, "transformer.1.logstd", c)
    z0 = mean_p + eps * torch.exp(logstd_p.float()) * temperature              # :268-271
    z_new = flow_reverse(sd, "transformer_flow", hp["transformer"]["flow"], z0, c)   # :272-275
    logits = _lin(sd, "token_predictor.linear", F.relu(_lin(sd, "token_spliter.linear", H)))   # :278-279
    tok = logits.argmax(-1, keepdim=True).float()
    return {"output": torch.cat([tok, z_new], -1), "kv": kvs, "logits": logits, "transformer_latent": H}
