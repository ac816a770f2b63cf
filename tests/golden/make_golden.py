"""Generate the golden fixtures of tests/golden/ by running the REAL reference (/root/reference).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
It instantiates the reference ``models.speech.lvtr.LVTR`` on a SMALL configuration of the same
architecture (so the fixture stays < 1 MB), injects the five RNG draws of ``LVTR.forward`` by patching
``torch.randn_like / torch.rand / torch.randint`` (the reference calls them through the torch namespace:
linear/layers.py:114, lvtr.py:334, ddpm.py:372,354), runs forward + backward with the loss assembly of
trainers/speech/lvtr.py:122-130, then a prefill + 3 cached ``LVTR.step`` calls with injected prior eps
and ``torch.multinomial`` patched to argmax (greedy), and stores inputs, weights, outputs and gradients.
"""
import copy
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_shim import import_reference  # noqa: E402

warnings.filterwarnings("ignore")


def small_config(n_mels: int = 16) -> dict:
    norm = {"identifier": "InstanceNorm", "eps": 1e-6}
    return {
        "tokens": {"embedding_dim": 64, "vocab_size": 50},
        "latent_dim": 4,
        "utterance_encoder": {
            "num_layers": 3, "resample_rates": [-2, -2, -2], "resample_ksize": [4, 4, 4], "init_channel": 8,
            "out_channels": [8, 16, 16], "layer": {"norm": dict(norm), "activation": {"identifier": "ReLU"}},
            "embedding_dim": 8},
        "encoder": {
            "identifier": "BottleNeckResNet", "num_layers": 2, "resample_rates": [1, 1], "resample_ksize": [1, 1],
            "init_channel": 16, "out_channels": [16, 16], "hidden_channels": [32, 32], "final_norm": True,
            "layer": {"causal_padding": True, "kernel_size": 7, "norm": dict(norm),
                      "activation": {"identifier": "ReLU"}}},
        "decoder": {
            "diffusion": {"timesteps": 1000, "loss_type": "l1", "input_scale": 5.0, "objective": "pred_noise",
                          "clamp_range": [-3.0, 1.2], "ddim_sampling_eta": 1.0,
                          "beta_schedule": {"identifier": "cosine"}, "identifier": "ConditionalBottleNeckUNet"},
            "cond_unet": {
                "unet": {"condition_dim": 8, "num_layers": 4, "resample_rates": [1, 1, 1, 1],
                         "resample_ksize": [1, 1, 1, 1], "init_channel": 16, "out_channels": [16, 16, 16, 16],
                         "hidden_channels": [32, 32, 32, 32], "conditional": [False, True, True, False],
                         "skip_connection": [None, None, 1, 0], "connection_type": "concat", "final_norm": True,
                         "layer": {"causal_padding": True, "kernel_size": 7, "norm": dict(norm),
                                   "activation": {"identifier": "SiLU"}, "condition_type": "concat"},
                         "upward_layer": {"boundary": 2, "future_padding": True, "kernel_size": 7, "norm": dict(norm),
                                          "activation": {"identifier": "SiLU"}, "condition_type": "concat"}},
                "time_embedding": {"dim": 16, "maxpos": 1000, "activation": {"identifier": "SiLU"}}}},
        "transformer": {
            "bias": False, "rpe": {"identifier": "ALiBi", "maxpos": 64}, "num_layers": 2,
            "layer": {"ffd_size": 128, "dim": 128, "norm": {"identifier": "RMSNorm", "eps": 1e-6},
                      "activation": {"identifier": "GELU"}, "self_attn": {"nheads": 2, "causal": True}},
            "flow": {"num_layers": 4, "conditional": True,
                     "layer": {"hidden_dim": 64, "activation": {"identifier": "GELU"}, "mean_only": False,
                               "scale_range": [0.5, 2.0], "norm": {"identifier": "LayerNorm", "eps": 1e-6}}}},
    }


class PatchedRNG:
    """Feed a fixed list of tensors to the reference's RNG calls, in call order."""

    def __init__(self, **queues):
        self.queues = {k: list(v) for k, v in queues.items()}
        self.saved = {}

    def _pop(self, name):
        return self.queues[name].pop(0)

    def __enter__(self):
        self.saved = dict(randn_like=torch.randn_like, rand=torch.rand, randint=torch.randint,
                          multinomial=torch.multinomial)
        torch.randn_like = lambda t, *a, **k: self._pop("randn_like").to(t.dtype)
        torch.rand = lambda *a, **k: self._pop("rand")
        torch.randint = lambda *a, **k: self._pop("randint")
        torch.multinomial = lambda probs, n, *a, **k: probs.argmax(-1, keepdim=True)
        return self

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            setattr(torch, k, v)


def main():
    LVTR, Hparams, TensorMask = import_reference()
    torch.manual_seed(20260417)
    n_mels = 16
    cfg = small_config(n_mels)
    model = LVTR(Hparams.from_json(__import__("json").dumps(cfg)), input_dim=n_mels)
    # perturb every parameter a little so biases / norm scales are exercised (default init leaves many at 0 / 1)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    # the 11 diffusion-schedule buffers that the training loss never reads are deterministic functions of the
    # config; leave them out of the fixture to keep it small
    used_buffers = ("decoder.sqrt_alphas_cumprod", "decoder.sqrt_one_minus_alphas_cumprod")
    buffer_names = {"decoder." + n for n, _ in model.decoder.named_buffers(recurse=False)}
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()
          if k not in buffer_names or k in used_buffers}

    B, T, Tu = 2, 24, 16
    lengths = torch.tensor([24, 17])
    utt_lengths = torch.tensor([16, 11])
    tokens = torch.randint(0, 50, (B, T), generator=g)
    mel = torch.randn(B, T, n_mels, generator=g)
    utt = torch.randn(B, Tu, n_mels, generator=g)
    mask = torch.arange(T)[None, :] < lengths[:, None]
    utt_mask = torch.arange(Tu)[None, :] < utt_lengths[:, None]
    x = torch.cat([tokens[..., None].float(), mel], -1)
    rng = {"eps_q": torch.randn(B, T, 4, generator=g), "init_state": torch.rand(B, 1, 64, generator=g) * 2 - 1,
           "eps_p": torch.randn(B, T, 4, generator=g), "diff_t": torch.randint(0, 1000, (B,), generator=g),
           "diff_noise": torch.randn(B, T, n_mels, generator=g)}

    # ---------------- forward + backward (trainers/speech/lvtr.py:111-131), kw = 0.04
    kw = 0.04
    with PatchedRNG(randn_like=[rng["eps_q"], rng["eps_p"], rng["diff_noise"]],
                    rand=[(rng["init_state"] + 1) / 2], randint=[rng["diff_t"]]):
        out = model(TensorMask(x, mask), utterance=TensorMask(utt, utt_mask))
    a = out["log_q"].flatten().apply_mask().value
    b = out["log_p"].flatten().apply_mask().value
    kld = (a - b).mean(-1).sum(-1).sum()
    loss = out["decoder_output"] * 1.0 + kld * kw + out["ce_loss"] * 0.5 * kw
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    assert all(gr is not None for gr in grads.values())
    logits = model.token_predictor(model.token_spliter(out["transformer_latent"])).value.detach()
    fwd = {"loss": loss.detach(), "kld": kld.detach(), "rec": out["decoder_output"].detach(),
           "ce": out["ce_loss"].detach(), "log_p": out["log_p"].value.detach(), "log_q": out["log_q"].value.detach(),
           "transformer_latent": out["transformer_latent"].value.detach(), "sample_q": out["sample_q"].value.detach(),
           "logits": logits, "u_c": out["u_c"].detach(), "kw": torch.tensor(kw)}

    # ---------------- cached decode: prefill over a 6-frame prompt (+BOS) then 3 single steps, greedy, eps injected
    Tp, nsteps = 6, 4
    prompt = torch.cat([tokens[:, :Tp, None].float(), torch.randn(B, Tp, 4, generator=g)], -1)
    s0 = torch.rand(B, 1, 64, generator=g) * 2 - 1
    eps_steps = [torch.randn(B, Tp + 1, 4, generator=g)] + [torch.randn(B, 1, 4, generator=g) for _ in range(nsteps - 1)]
    dec = {"prompt": prompt, "init_state": s0, "eps": eps_steps, "outputs": [], "latents": [], "logits": []}
    state, kv = prompt, None
    model.eval()
    with torch.no_grad():
        for i in range(nsteps):
            with PatchedRNG(randn_like=[eps_steps[i]], rand=[(s0 + 1) / 2], randint=[]):
                o = model.step(state, past_kv=kv, temperature=0.85, token_temperature=0.85,
                               push_init_state=(i == 0))
            lat = o["transformer_latent"].value
            dec["latents"].append(lat.clone())
            dec["logits"].append(model.token_predictor(model.token_spliter(o["transformer_latent"])).value.clone())
            dec["outputs"].append(o["output"].clone())
            kv = o["kv"]
            state = o["output"][:, -1:]
    dec["final_kv_key0"] = kv[0]["key"].clone()

    # ---------------- diffusion decoding of generated frames (lvtr.py:288-306 → ddpm.py:284-321), DDIM with 6
    # sampling steps; the start noise (torch.randn) and the per-step noises (torch.randn_like) are injected
    nsamp = 6
    model.decoder.sampling_timesteps = nsamp
    frames = torch.cat([tokens[..., None].float(), torch.randn(B, T, 4, generator=g)], -1)
    start = torch.randn(B, T, n_mels, generator=g)
    step_noise = [torch.randn(B, T, n_mels, generator=g) for _ in range(nsamp - 1)]
    saved_randn = torch.randn
    torch.randn = lambda *a, **k: start.clone()
    try:
        with torch.no_grad(), PatchedRNG(randn_like=step_noise, rand=[], randint=[]):
            mel_out = model.decode(TensorMask(frames, mask), u_c=out["u_c"].detach())
    finally:
        torch.randn = saved_randn
    ddim = {"frames": frames, "u_c": out["u_c"].detach().clone(), "start": start, "noise": step_noise,
            "steps": nsamp, "output": mel_out.value.detach().clone()}

    # ---------------- likelihood scoring (lvtr.py:337-388), temperature-0 posterior, injected BOS state
    s0l = torch.rand(B, 1, 64, generator=g) * 2 - 1
    with torch.no_grad(), PatchedRNG(randn_like=[rng["eps_q"], rng["eps_p"]], rand=[(s0l + 1) / 2], randint=[]):
        ll = model.likelihood(TensorMask(x, mask), temperature=0.0)
    like = {"init_state": s0l, "value": ll.detach().clone()}

    fixture = {"config": cfg, "n_mels": n_mels, "state_dict": sd, "ddim": ddim, "likelihood": like,
               "inputs": {"x": x, "mask": mask, "utterance": utt, "utt_mask": utt_mask, **rng},
               "forward": fwd, "grads": grads, "decode": dec,
               "note": "generated by tests/golden/make_golden.py from the unmodified reference (torch %s)" % torch.__version__}
    path = os.path.join(HERE, "lvtr_small.pt")
    torch.save(fixture, path)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024),
          "params:", sum(v.numel() for v in sd.values()))
    print({k: float(v) for k, v in fwd.items() if v.dim() == 0})


if __name__ == "__main__":
    main()
