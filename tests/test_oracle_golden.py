"""The oracle (oracle/lvtr_oracle.py) against the golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only; this is what pins the oracle."""
import torch

from oracle import lvtr_oracle as O

TOL = dict(rtol=2e-5, atol=2e-6)


def _inputs(g):
    i = g["inputs"]
    rng = {k: i[k] for k in ("eps_q", "init_state", "eps_p", "diff_t", "diff_noise")}
    return i["x"], i["mask"], i["utterance"], i["utt_mask"], rng


def test_forward_matches_reference(golden):
    x, mask, utt, um, rng = _inputs(golden)
    out = O.lvtr_forward(golden["state_dict"], golden["config"], x, mask, utt, um, rng)
    f = golden["forward"]
    for name, key in (("decoder_output", "rec"), ("kld", "kld"), ("ce_loss", "ce"), ("log_p", "log_p"),
                      ("log_q", "log_q"), ("transformer_latent", "transformer_latent"), ("sample_q", "sample_q"),
                      ("logits", "logits"), ("u_c", "u_c")):
        torch.testing.assert_close(out[name], f[key], **TOL, msg=lambda m, n=name: f"{n}: {m}")
    torch.testing.assert_close(O.total_loss(out, float(f["kw"])), f["loss"], **TOL)


def test_gradients_match_reference(golden):
    x, mask, utt, um, rng = _inputs(golden)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in golden["state_dict"].items()}
    out = O.lvtr_forward(sd, golden["config"], x, mask, utt, um, rng)
    O.total_loss(out, float(golden["forward"]["kw"])).backward()
    assert set(golden["grads"]) <= set(sd)
    for k, gref in golden["grads"].items():
        assert sd[k].grad is not None, k
        torch.testing.assert_close(sd[k].grad, gref, rtol=2e-4, atol=2e-6, msg=lambda m, n=k: f"grad {n}: {m}")


def test_padded_rows_are_zero(golden):
    x, mask, utt, um, rng = _inputs(golden)
    out = O.lvtr_forward(golden["state_dict"], golden["config"], x, mask, utt, um, rng)
    pad = ~mask
    assert pad.any()
    for k in ("transformer_latent", "log_p", "log_q", "sample_q"):
        assert float(out[k][pad].abs().max()) == 0.0, k


def test_cached_decode_matches_reference(golden):
    d = golden["decode"]
    sd, cfg = golden["state_dict"], golden["config"]
    state, kv = d["prompt"], None
    for i, eps in enumerate(d["eps"]):
        o = O.lvtr_step(sd, cfg, state, kv, eps, temperature=0.85, push_init_state=(i == 0),
                        init_state=d["init_state"])
        torch.testing.assert_close(o["transformer_latent"], d["latents"][i], **TOL)
        torch.testing.assert_close(o["logits"], d["logits"][i], rtol=2e-5, atol=2e-5)
        assert torch.equal(o["output"][..., 0], d["outputs"][i][..., 0]), "greedy token ids must be bit-exact"
        torch.testing.assert_close(o["output"][..., 1:], d["outputs"][i][..., 1:], rtol=2e-5, atol=2e-5)
        kv = o["kv"]
        state = o["output"][:, -1:]
    torch.testing.assert_close(kv[0]["key"], d["final_kv_key0"], **TOL)


def test_ddim_decode_matches_reference(golden):
    """LVTR.decode → ddim_sample (lvtr.py:288-306, ddpm.py:284-321), 6 sampling steps, injected noises."""
    d, i = golden["ddim"], golden["inputs"]
    out = O.lvtr_decode(golden["state_dict"], golden["config"], d["frames"], i["mask"], d["u_c"], d["start"], d["noise"],
                        d["steps"])
    torch.testing.assert_close(out, d["output"], rtol=1e-4, atol=1e-4)
    assert float(out[~i["mask"]].abs().max()) == 0.0


def test_likelihood_matches_reference(golden):
    """LVTR.likelihood (lvtr.py:337-388), temperature-0 posterior."""
    i, l = golden["inputs"], golden["likelihood"]
    val = O.lvtr_likelihood(golden["state_dict"], golden["config"], i["x"], i["mask"], l["init_state"])
    torch.testing.assert_close(val, l["value"], rtol=2e-5, atol=2e-5)
