"""HiFi-GAN generator (vae_gslm_b200/models/vocoder/hifigan.py) against the REAL reference generator
(models/vocoder/hfgan.py) with the same weight-normed checkpoint, on CPU; plus the checkpoint-format handling."""
import os
import sys

import pytest
import torch

from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.models.vocoder.hifigan import Generator, HiFiGAN, fold_weight_norm
from vae_gslm_b200.utils.tensormask import TensorMask

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from ref_shim import import_reference, reference_available  # noqa: E402

GEN = {"weight_norm": True, "upsample_rates": [5, 4, 2], "upsample_kernel_sizes": [10, 8, 4],
       "upsample_initial_channel": 32, "resblock_kernel_sizes": [3, 7], "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5]],
       "in_channels": 12, "kernel_size": 7}


def _mel(B=3, T=23, C=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    return TensorMask.fromlength(torch.randn(B, T, C, generator=g), torch.tensor([T, 11, 1])).apply_mask()


@pytest.mark.skipif(not reference_available(), reason="the reference tree is only present in the build container")
def test_generator_matches_reference(tmp_path):
    _, RefHp, RefTM = import_reference()
    from models.vocoder.hfgan import Generator as RefGenerator
    from models.vocoder.vocoder import HiFiGAN as RefHiFiGAN
    import json
    torch.manual_seed(0)
    ref = RefGenerator(RefHp.from_json(json.dumps(GEN)) if hasattr(RefHp, "from_json") else RefHp.from_dict(GEN))
    with torch.no_grad():                       # make the weight-norm gains non-trivial
        for n, p in ref.named_parameters():
            if n.endswith("original0"):
                p.mul_(1.0 + 0.5 * torch.rand_like(p))
            if n.endswith("original1") or n.endswith("bias"):
                p.copy_(0.3 * torch.randn_like(p))
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    assert any("parametrizations" in k or k.endswith("weight_g") for k in sd)
    ours = Generator(Hparams.from_dict(GEN))
    ours.load_reference_state_dict(sd)
    x = _mel()
    with torch.no_grad():
        want = ref(RefTM(x.value, x.mask))
    got = ours(x)
    assert torch.equal(got.mask, want.mask) and got.value.shape == (3, 23 * 40)
    assert float((got.value - want.value).abs().max()) < 1e-5
    # after the reference's own remove_weight_norm the plain checkpoint loads directly and agrees too
    ref.remove_weight_norm()
    plain = Generator(Hparams.from_dict(GEN))
    plain.load_reference_state_dict(ref.state_dict())
    assert float((plain(x).value - want.value).abs().max()) < 1e-5
    # wrapper: rescale + masking, through from_pretrained on the reference's on-disk layout
    full = {"feature": {"sample_rate": 16000, "n_mels": 12}, "model": {"generator": GEN}}
    import yaml
    (tmp_path / "hp.yaml").write_text(yaml.safe_dump(full))
    torch.save(sd, tmp_path / "last-cpt.ckpt")
    rs = {"mean": -1.5, "std": 2.0}
    mine = HiFiGAN.from_pretrained(str(tmp_path), hp_rescale=Hparams.from_dict(rs))
    theirs = RefHiFiGAN.from_pretrained(str(tmp_path), hp_rescale=RefHp(**rs))
    with torch.no_grad():
        w = theirs.decode(RefTM(x.value, x.mask))
    m = mine.decode(x)
    assert float((m.value - w.value).abs().max()) < 1e-5 and float(m.value[~m.mask].abs().max()) == 0.0


def test_fold_weight_norm_formats_and_shapes():
    torch.manual_seed(1)
    v, g = torch.randn(6, 4, 3), torch.rand(6, 1, 1) + 0.5
    want = g * v / v.flatten(1).norm(dim=1).view(6, 1, 1)
    new = fold_weight_norm({"c.parametrizations.weight.original0": g, "c.parametrizations.weight.original1": v, "c.bias": torch.zeros(6)})
    old = fold_weight_norm({"c.weight_g": g, "c.weight_v": v, "c.bias": torch.zeros(6)})
    assert set(new) == set(old) == {"c.weight", "c.bias"}
    assert torch.allclose(new["c.weight"], want) and torch.allclose(old["c.weight"], want)
    gen = Generator(Hparams.from_dict(GEN))
    out = gen(_mel())
    assert out.value.shape == (3, 23 * 40) and out.length.tolist() == [920, 440, 40] and float(out.value.abs().max()) <= 1.0
