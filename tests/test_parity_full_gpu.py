"""Parity of the CUDA path at the FULL configuration (d = 1024, 16 layers, 226,957,564 parameters) and at BASELINE.json's
shapes against the oracle run on the same GPU with the same weights, inputs and injected draws:

* forward + backward at B=8,T=640 (the repo config), B=8,T=1000 (configs[1]) and B=1,T=3000 (configs[4]) — fp32 parity
  mode within 1e-4 (5e-4 Frobenius on gradients), bf16 mode within 2e-2 on loss / logits and, for the gradients, a
  per-parameter comparison of THREE quantities: ours vs the fp32 oracle, the oracle under torch.autocast(bf16) vs the fp32
  oracle (what stock PyTorch bf16 does to the same gradient) and ours vs the autocast oracle directly.  The table is written
  to gpurun_out/parity_bf16_<shape>.json (committed as profiles/r02_parity_bf16.md); the bound asserted here and in
  __graft_entry__.smoke() is BF16_GRAD_BOUND below;
* cached generation at the full configuration: prefill of 150 frames + BOS, then 32 cached single-frame steps, teacher-forced
  with the oracle's outputs: fp32 greedy token ids bit-exact, bf16 engines (persistent step kernel, kernel-per-linear engine,
  layer-by-layer) against the ORACLE's logits (not against each other);
* ARTRSampler.__call__ end to end (prompt → prefill → cached steps incl. the CUDA-graph replay → DDIM decode)."""
import copy
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import lvtr_oracle as O                                     # noqa: E402  (checker only)
from vae_gslm_b200 import ops                                           # noqa: E402
from vae_gslm_b200.hparams.hp import Hparams                            # noqa: E402
from vae_gslm_b200.models.speech.lvtr import LVTR                       # noqa: E402
from vae_gslm_b200.trainers.speech.lvtr import assemble_loss            # noqa: E402
from vae_gslm_b200.utils.tensormask import TensorMask                   # noqa: E402

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "vae_gslm_b200", "configs", "train", "speech", "vae-gslm.yaml")

# bf16 gradient bound: O.bf16_grad_within_bound (oracle/lvtr_oracle.py) — per tensor within 2e-2 of the fp32 oracle gradient or
# no noisier than 2.5 x the oracle under torch.autocast(bf16); the same function is asserted by __graft_entry__.smoke()
BF16_GRAD_BOUND = O.BF16_GRAD_BOUND


def frob_rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def max_rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


@pytest.fixture(autouse=True)
def _exact_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ops.GEMM_BACKEND = ops.GEMM_AUTO
    yield
    ops.GEMM_BACKEND = ops.GEMM_AUTO


def _relu_flip_only(mine, ref, max_rows=4):
    """A gradient tensor behind a ReLU may differ in a few output features only: with ~1e7 pre-activations per step a
    handful sit within fp32 rounding of zero and take the other branch on the two sides (observed: ONE element of
    token_spliter at B=8, T=640 — 20 % of that feature's row, 7.7e-4 of the tensor, every other row at 3e-6).  That is a
    property of the function, not of the implementation: accept it when all but `max_rows` rows agree to 3e-4."""
    m2, r2 = mine.float().reshape(mine.shape[0], -1), ref.float().reshape(ref.shape[0], -1)
    row = (m2 - r2).norm(dim=1) / (r2.norm(dim=1) + 1e-20)
    return int((row > 3e-4).sum()) <= max_rows


def full_model(seed=0):
    from vae_gslm_b200.training_lib.trainer import init_weights
    torch.manual_seed(seed)
    model = LVTR(Hparams.from_yamlfile(CFG).model, input_dim=80)
    model.apply(init_weights)
    # the reference's init zeroes every bias: give them values so that bias paths are exercised by the comparison
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return model.to(DEV), Hparams.from_yamlfile(CFG).model.to_dict()


def synthetic(B, T, ragged=True):
    g = torch.Generator(device="cpu").manual_seed(1234)
    lengths = torch.randint(T // 2, T + 1, (B,), generator=g) if ragged else torch.full((B,), T)
    lengths[0] = T
    Tu = 150
    ul = torch.randint(100, Tu + 1, (B,), generator=g)
    tokens = torch.randint(0, 200, (B, T), generator=g)
    mel = torch.randn(B, T, 80, generator=g)
    batch = {"x": torch.cat([tokens[..., None].float(), mel], -1), "mask": torch.arange(T)[None] < lengths[:, None],
             "utterance": torch.randn(B, Tu, 80, generator=g), "utt_mask": torch.arange(Tu)[None] < ul[:, None]}
    g2 = torch.Generator(device="cpu").manual_seed(4321)
    rng = {"eps_q": torch.randn(B, T, 4, generator=g2), "init_state": torch.rand(B, 1, 64, generator=g2) * 2 - 1,
           "eps_p": torch.randn(B, T, 4, generator=g2), "diff_t": torch.randint(0, 1000, (B,), generator=g2),
           "diff_noise": torch.randn(B, T, 80, generator=g2)}
    return {k: v.to(DEV) for k, v in batch.items()}, {k: v.to(DEV) for k, v in rng.items()}


def oracle_run(sd_src, cfg, batch, rng, kw, autocast):
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k in sd_src["_param_names"])
          for k, v in sd_src.items() if k != "_param_names"}
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with ctx:
        out = O.lvtr_forward(sd, cfg, batch["x"], batch["mask"], batch["utterance"], batch["utt_mask"], rng)
        loss = O.total_loss(out, kw)
    loss.backward()
    res = {"loss": float(loss), "kld": float(out["kld"]), "ce": float(out["ce_loss"]), "rec": float(out["decoder_output"]),
           "logits": out["logits"].detach().float(), "latent": out["transformer_latent"].detach().float(),
           "grads": {k: v.grad.detach().clone() for k, v in sd.items() if v.requires_grad}}
    del out, loss, sd
    torch.cuda.empty_cache()
    return res


@pytest.mark.parametrize("B,T", [(8, 640), (8, 1000), (1, 3000)])
def test_full_config_forward_backward_at_baseline_shapes(B, T):
    model, cfg = full_model()
    batch, rng = synthetic(B, T)
    kw = 0.04
    sd_src = dict(model.state_dict())
    sd_src["_param_names"] = {n for n, _ in model.named_parameters()}
    ref = oracle_run(sd_src, cfg, batch, rng, kw, autocast=False)
    ref16 = oracle_run(sd_src, cfg, batch, rng, kw, autocast=True)
    table = {}
    for mode in ("fp32", "bf16"):
        model.set_compute_dtype(torch.float32 if mode == "fp32" else torch.bfloat16)
        model.zero_grad(set_to_none=True)
        out = model(TensorMask(batch["x"], batch["mask"]), utterance=TensorMask(batch["utterance"], batch["utt_mask"]), **rng)
        terms = assemble_loss(out, kld_weight=kw)
        terms["loss"].backward()
        tol = 1e-4 if mode == "fp32" else 2e-2
        for name, mine, r in (("loss", terms["loss"], ref["loss"]), ("kld", terms["kld"], ref["kld"]),
                              ("ce", terms["token_kld"], ref["ce"]), ("rec", terms["rec_loss"], ref["rec"])):
            assert abs(float(mine) - r) / abs(r) < tol, (mode, name, float(mine), r)
        assert max_rel(out["logits"].value, ref["logits"]) < tol, mode
        assert max_rel(out["transformer_latent"].value, ref["latent"]) < (tol if mode == "fp32" else 4e-2), mode
        pad = ~batch["mask"]
        if bool(pad.any()):
            assert float(out["transformer_latent"].value[pad].abs().max()) == 0.0
        bad = []
        for name, p in model.named_parameters():
            e = frob_rel(p.grad, ref["grads"][name])
            if mode == "fp32":
                # 5e-4: the checker is itself an fp32 cuBLAS computation with its own summation order; at 8000 frames the
                # two sides differ by up to 4e-4 on gradients that are long scatter-sums (token_embedding), median 2e-5
                if e > 5e-4 and not _relu_flip_only(p.grad, ref["grads"][name]):
                    bad.append((name, e))
            else:
                e_ac = frob_rel(ref16["grads"][name], ref["grads"][name])
                e_direct = frob_rel(p.grad, ref16["grads"][name])
                table[name] = {"numel": p.numel(), "ours_vs_fp32": e, "autocast_vs_fp32": e_ac, "ours_vs_autocast": e_direct}
                if not O.bf16_grad_within_bound(e, e_ac):
                    bad.append((name, round(e, 5), round(e_ac, 5)))
        assert not bad, (mode, bad[:12])
        del out, terms
    # whole-gradient (all parameters concatenated) error: the number the optimizer sees
    num = sum(float((p.grad.float() - ref["grads"][n]).pow(2).sum()) for n, p in model.named_parameters())
    den = sum(float(ref["grads"][n].pow(2).sum()) for n, _ in model.named_parameters())
    num_ac = sum(float((ref16["grads"][n] - ref["grads"][n]).pow(2).sum()) for n, _ in model.named_parameters())
    summary = {"B": B, "T": T, "whole_gradient_ours_vs_fp32": (num / den) ** 0.5,
               "whole_gradient_autocast_vs_fp32": (num_ac / den) ** 0.5,
               "worst_tensor_ours_vs_fp32": max(v["ours_vs_fp32"] for v in table.values()),
               "n_tensors_above_2e-2": sum(v["ours_vs_fp32"] > 2e-2 for v in table.values()),
               "n_tensors_autocast_above_2e-2": sum(v["autocast_vs_fp32"] > 2e-2 for v in table.values())}
    print("bf16 gradient summary", json.dumps(summary))
    assert summary["whole_gradient_ours_vs_fp32"] < BF16_GRAD_BOUND
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"parity_bf16_B{B}_T{T}.json"), "w") as f:
            json.dump({"summary": summary, "tensors": table}, f, indent=0)
    except OSError:
        pass


# ----------------------------------------------------------------------------------------- cached generation, full size
def _oracle_generation(model, cfg, B, prompt_len, steps, seed=5):
    """the oracle's own greedy continuation (fp32, eps injected): prompt frames, per-step outputs / logits / latents"""
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    g = torch.Generator(device="cpu").manual_seed(seed)
    prompt = torch.cat([torch.randint(0, 200, (B, prompt_len, 1), generator=g).float(),
                        torch.randn(B, prompt_len, 4, generator=g)], -1).to(DEV)
    init_state = (torch.rand(B, 1, 64, generator=g) * 2 - 1).to(DEV)
    eps = [torch.randn(B, prompt_len + 1 if i == 0 else 1, 4, generator=g).to(DEV) for i in range(steps + 1)]
    outs, state, kv = [], prompt, None
    with torch.no_grad():
        for i in range(steps + 1):
            o = O.lvtr_step(sd, cfg, state, kv, eps[i], temperature=0.85, push_init_state=(i == 0), init_state=init_state)
            kv = o["kv"]
            state = o["output"][:, -1:]
            outs.append({"output": o["output"][:, -1:], "logits": o["logits"][:, -1:], "latent": o["transformer_latent"][:, -1:]})
    return prompt, init_state, eps, outs


@pytest.mark.parametrize("mode", ["fp32", "bf16-step", "bf16-linear", "bf16-linear-skinny", "bf16-layerwise",
                                  "bf16-layerwise-b80"])
def test_full_config_cached_generation_against_oracle(mode):
    # b80: 80 sequences — the layer-by-layer step whose FFN takes the split-K route of ops.ffn (65..256 rows)
    B, P, S = (80, 150, 6) if mode.endswith("b80") else (2, 150, 32)
    model, cfg = full_model()
    model.eval()
    prompt, init_state, eps, ref = _oracle_generation(model, cfg, B, P, S)
    fp32 = mode == "fp32"
    model.set_compute_dtype(torch.float32 if fp32 else torch.bfloat16)
    if not fp32:
        kind = mode.split("-")[1]
        model.use_decode_engine = kind != "layerwise"
        model.decode_engine_kind = kind
        model.decode_engine_skinny = mode.endswith("skinny")       # the engine's linears on vg_skinny_linear (folded RMSNorm)
    model.transformer[0].cache_len_hint = P + S + 8
    state, kv = prompt, None
    top2_gap = []
    for i in range(S + 1):
        o = model.step(state, past_kv=kv, temperature=0.85, push_init_state=(i == 0), eps=eps[i], greedy=True,
                       init_state=init_state, return_logits=True)
        kv = o["kv"]
        lg = o["logits"][:, -1:].float()
        r = ref[i]
        if fp32:
            assert max_rel(lg, r["logits"]) < 1e-4, i
            assert max_rel(o["transformer_latent"].value[:, -1:], r["latent"]) < 1e-4, i
            assert torch.equal(o["output"][:, -1:, 0], r["output"][..., 0]), f"greedy token ids differ at step {i}"
            assert max_rel(o["output"][:, -1:, 1:], r["output"][..., 1:]) < 1e-4, i
        else:
            assert max_rel(lg, r["logits"]) < 3e-2, (mode, i, max_rel(lg, r["logits"]))
            assert max_rel(o["transformer_latent"].value[:, -1:], r["latent"]) < 4e-2, (mode, i)
            # the greedy token may only differ where the oracle's own top-2 logits are closer than the bf16 error
            t2 = r["logits"].float().topk(2, -1).values
            gap = (t2[..., 0] - t2[..., 1]).reshape(-1)
            same = (o["output"][:, -1:, 0] == r["output"][..., 0]).reshape(-1)
            top2_gap += [float(x) for x in gap[~same]]
        state = r["output"]                                   # teacher-force the oracle's own output
    if not fp32:
        scale = float(ref[0]["logits"].abs().max())
        assert all(gp < 3e-2 * scale for gp in top2_gap), top2_gap
        if kind != "layerwise":
            eng = model.__dict__["_decode_engines"][B][1]
            assert type(eng).__name__ == ("DecodeStepEngine" if mode == "bf16-step" else "DecodeEngine")
            assert getattr(eng, "skinny", False) == mode.endswith("skinny")
    assert kv[0].cache.length == P + 1 + S


def test_artr_sampler_matches_oracle_loop(golden):
    """ARTRSampler.__call__ (trainers/speech/sampler.py:17-72): encode the prompt, prefill, cached steps (eager, then the
    CUDA-graph replay from the fourth step on) and DDIM decode.  Deterministic settings (encoder temperature 0, prior
    temperature 0, greedy tokens, fixed BOS state) so that the frames can be held to the oracle's own free-running loop."""
    from vae_gslm_b200.trainers.speech.sampler import ARTRSampler
    cfg = golden["config"]
    model = LVTR(Hparams.from_dict(copy.deepcopy(cfg)), input_dim=golden["n_mels"])
    model.load_state_dict(golden["state_dict"], strict=False)
    model = model.to(DEV).set_compute_dtype(torch.float32).eval()
    sd = {k: v.to(DEV) for k, v in golden["state_dict"].items()}
    i = golden["inputs"]
    B, P, length = 2, 24, 12
    prior = i["x"][:B, :P].to(DEV)                            # [B, P, 1 + n_mels]: token id + mel frames of the prompt
    s0 = (torch.rand(B, 1, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1).to(DEV)
    model.initial_state = lambda bsize, device=None, nfeat=None: s0
    model.decoder.sampling_timesteps, model.decoder.ddim_sampling_eta = 4, 0.0
    out = ARTRSampler(model)(length, prior, temperature=0.0, token_temperature=1.0, encoder_temperature=0.0, greedy=True)
    frames = out["frames"]
    assert frames.shape == (B, P + length, 5)
    # oracle: posterior mean of the prompt, then the free-running greedy loop
    mask = torch.ones(B, P, dtype=torch.bool, device=DEV)
    h_enc = O.bottleneck_resnet(sd, "encoder.0", cfg["encoder"], prior[..., 1:], mask)
    z = O._lin(sd, "encoder.1.mean", h_enc)
    state, kv = torch.cat([prior[..., :1], z], -1), None
    want = [state]
    with torch.no_grad():
        for k in range(length):
            n_new = state.shape[1] + (1 if k == 0 else 0)
            o = O.lvtr_step(sd, cfg, state, kv, torch.zeros(B, n_new, 4, device=DEV), temperature=0.0,
                            push_init_state=(k == 0), init_state=s0)
            kv, state = o["kv"], o["output"][:, -1:]
            want.append(state)
    want = torch.cat(want, 1)
    assert torch.equal(frames[..., 0], want[..., 0]), "sampled token ids differ from the oracle's greedy loop"
    assert max_rel(frames[..., 1:], want[..., 1:]) < 1e-3
    mel = out["output"].value
    assert mel.shape == (B, P + length, golden["n_mels"]) and bool(torch.isfinite(mel).all())
    # bf16 product path (persistent step kernel under the CUDA graph): runs and stays close to the fp32 frames
    model.set_compute_dtype(torch.bfloat16)
    out16 = ARTRSampler(model)(length, prior, temperature=0.0, token_temperature=1.0, encoder_temperature=0.0, greedy=True,
                               decode=False)
    assert out16["frames"].shape == frames.shape
    assert "_decode_engines" in model.__dict__
    assert max_rel(out16["frames"][:, :P + 2, 1:], want[:, :P + 2, 1:]) < 6e-2
