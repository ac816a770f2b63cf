"""Host-side logic of the hot path against the REAL reference (differential, CPU): the TensorMask carrier
(utils/tensormask.py), padding helpers (utils/helpers.py:138-150), Hparams (hparams/hp.py), the KL-weight schedule
(trainers/speech/lvtr.py:104-110) and the learning-rate schedule of the recipe (training_lib/optimizer.py:58-107).
The reference tree only exists in the build container; there the differential tests run, elsewhere only the
closed-form checks do."""
import os
import sys

import pytest
import torch

from vae_gslm_b200.arena import cosine_lr
from vae_gslm_b200.hparams.hp import Hparams
from vae_gslm_b200.trainers.speech.lvtr import kld_weight_at
from vae_gslm_b200.utils.helpers import get_padding, make_padding_mask
from vae_gslm_b200.utils.tensormask import TensorMask

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from ref_shim import REFERENCE_ROOT, import_reference, reference_available  # noqa: E402

needs_ref = pytest.mark.skipif(not reference_available(), reason="the reference tree is only present in the build container")


def _pair(seed=0, B=3, T=9, C=6):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, T, C, generator=g)
    length = torch.tensor([T, 5, 1])[:B]
    return v, length


def _same(a, b):
    assert torch.equal(a.value, b.value) and torch.equal(a.mask, b.mask)


@needs_ref
def test_tensormask_matches_reference():
    _, _, RefTM = import_reference()
    v, length = _pair()
    ours, ref = TensorMask.fromlength(v, length), RefTM.fromlength(v, length)
    _same(ours, ref)
    _same(ours.apply_mask(), ref.apply_mask())
    assert torch.equal(ours.length, ref.length) and len(ours) == len(ref) and ours.size(1) == ref.size(1)
    first = torch.randn(3, 1, 6)
    _same(ours.push(first), ref.push(first))
    _same(ours.push(first).pop(), ref.push(first).pop())
    _same(ours.pop(2), ref.pop(2))
    a, b = ours.split(2)
    ra, rb = ref.split(2)
    _same(a, ra)
    _same(b, rb)
    _same(a.cat(b), ra.cat(rb))
    ids, rids = TensorMask(v[..., 0], ours.mask), RefTM(v[..., 0], ref.mask)
    _same(ids.expand(), rids.expand())
    _same(ids.expand().cat(ours), rids.expand().cat(ref))
    _same(ours.transpose(), ref.transpose())
    _same(ours.abs(), ref.abs())
    for o, r in zip(ours.tolist(), ref.tolist()):
        assert torch.equal(o, r)
    assert torch.equal(TensorMask.use_mask(v, ours.mask), RefTM.use_mask(v, ref.mask))
    assert torch.equal(TensorMask.resize_length(length, 0.5), RefTM.resize_length(length, 0.5))
    # default mask = all valid
    _same(TensorMask(v), RefTM(v))


def test_tensormask_invariants():
    v, length = _pair(1)
    tm = TensorMask.fromlength(v, length).apply_mask()
    assert bool((tm.value[~tm.mask] == 0).all()) and tm.length.tolist() == length.tolist()
    assert tm.row_mask_u8().dtype == torch.uint8 and tm.lengths_i32().dtype == torch.int32
    pushed = tm.push(torch.ones(3, 1, 6))
    assert pushed.value.shape[1] == 10 and pushed.length.tolist() == (length + 1).tolist()
    popped = pushed.pop()
    assert torch.equal(popped.value, pushed.value[:, :-1]) and popped.length.tolist() == length.tolist()


@needs_ref
def test_padding_helpers_match_reference():
    import_reference()
    from utils import helpers as ref_helpers
    for k, d, s in [(7, 1, 1), (3, 2, 1), (5, 1, 2), (1, 1, 1)]:
        for kw in ({}, {"causal": True}, {"future": True}):
            assert get_padding(k, d, s, **kw) == ref_helpers.get_padding(k, d, s, **kw)
    a = torch.rand(2, 5) > 0.3
    b = torch.rand(2, 7) > 0.3
    assert torch.equal(make_padding_mask(a, b), ref_helpers.make_padding_mask(a, b))


@needs_ref
def test_hparams_and_shipped_configs_match_reference():
    _, RefHp, _ = import_reference()
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("configs/train/speech/vae-gslm.yaml", "configs/infer/speech/vae-gslm.yaml"):
        ours = Hparams.from_yamlfile(os.path.join(here, "vae_gslm_b200", rel))
        ref = RefHp.from_yamlfile(os.path.join(REFERENCE_ROOT, rel))

        def plain(h):
            if hasattr(h, "__dict__"):
                return {k: plain(v) for k, v in vars(h).items()}
            if isinstance(h, (list, tuple)):
                return [plain(v) for v in h]
            return h

        # the shipped files keep only the hot path's keys: every value we ship must be the reference's, and the model
        # description (architecture = shapes of every kernel launch) must be complete
        def subset(a, b, path=""):
            if isinstance(a, dict):
                assert isinstance(b, dict), path
                for k, v in a.items():
                    assert k in b, f"{rel}: {path}/{k} is not a reference key"
                    subset(v, b[k], f"{path}/{k}")
            else:
                assert a == b, f"{rel}: {path}: {a!r} != {b!r}"

        subset(plain(ours), plain(ref))
        if "model" in plain(ours):
            assert plain(ours)["model"] == plain(ref)["model"], rel
    hp, rhp = Hparams.from_dict({"a": {"b": 1}, "c": [1, 2]}), RefHp(a=RefHp(b=1), c=[1, 2])
    assert hp.has("a") == rhp.has("a") and hp.get("zz", 7) == rhp.get("zz", 7) and hp.a.b == rhp.a.b
    hp.check_arg_in_hparams("a", "c")
    with pytest.raises(Exception):
        hp.check_arg_in_hparams("missing")
    with pytest.raises(Exception):
        rhp.check_arg_in_hparams("missing")


def _ref_kld(global_step, kld_scale, warmup_kld, zero_kld):
    """the reference's expression, restated (trainers/speech/lvtr.py:104-110)."""
    w = kld_scale
    if warmup_kld > 0 and ((global_step + 1) > zero_kld and (global_step + 1) <= warmup_kld):
        w = kld_scale * ((global_step - zero_kld) / warmup_kld)
    if zero_kld > 0 and global_step <= zero_kld:
        w = 0.0
    return w


def test_kld_weight_schedule():
    for warm, zero in [(0, 0), (30000, 0), (100, 10), (100, 100), (5, 50)]:
        for step in [0, 1, 9, 10, 11, 49, 50, 51, 99, 100, 101, 29999, 30000, 30001]:
            assert kld_weight_at(step, 0.04, warm, zero) == pytest.approx(_ref_kld(step, 0.04, warm, zero), rel=1e-12, abs=0), (warm, zero, step)


@needs_ref
def test_lr_schedule_matches_reference_scheduler():
    """flat 30 k steps then cosine to 5e-5 (configs/train/speech/vae-gslm.yaml:149-153), scaled down to 30 + 70 steps."""
    _, RefHp, _ = import_reference()
    from training_lib.optimizer import scheduler_map
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=5e-4)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sched, interval = scheduler_map(RefHp(identifier="cosine", min_lr=5e-5, flat_steps=30), opt, 100)
        assert interval == "step"
        for step in range(100):
            assert opt.param_groups[0]["lr"] == pytest.approx(cosine_lr(step, 5e-4, 5e-5, 30, 100), rel=1e-6), step
            opt.step()
            sched.step()


def test_lr_schedule_closed_form():
    assert cosine_lr(0, 5e-4, 5e-5, 30000, 1000000) == 5e-4 and cosine_lr(29999, 5e-4, 5e-5, 30000, 1000000) == 5e-4
    assert cosine_lr(1000000, 5e-4, 5e-5, 30000, 1000000) == pytest.approx(5e-5)
    mid = cosine_lr(30000 + 485000, 5e-4, 5e-5, 30000, 1000000)
    assert mid == pytest.approx((5e-4 + 5e-5) / 2)


def test_seeded_init_is_bit_identical_to_reference():
    """under torch.manual_seed(0) this repo's LVTR + init_weights produces, tensor for tensor and in
    the same state-dict order, the bytes of the REAL reference's LVTR + BaseTrainer.init_weights
    (training_lib/trainer.py:113-125) at the full configuration.  The reference side is the committed digest
    tests/golden/init_digest.json (made by tests/golden/make_init_digest.py in the build container)."""
    import hashlib
    import json
    from vae_gslm_b200.models.speech.lvtr import LVTR
    from vae_gslm_b200.training_lib.trainer import init_weights
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "init_digest.json")))
    hp = Hparams.from_yamlfile(os.path.join(root, "vae_gslm_b200", "configs", "train", "speech", "vae-gslm.yaml"))
    torch.manual_seed(ref["seed"])
    model = LVTR(hp.model, input_dim=80)
    model.apply(init_weights)
    assert sum(p.numel() for p in model.parameters()) == ref["n_params"] == 226957564
    sd = model.state_dict()
    assert list(sd.keys()) == [e["name"] for e in ref["entries"]]
    bad = []
    for e in ref["entries"]:
        t = sd[e["name"]].detach().cpu().contiguous()
        if list(t.shape) != e["shape"] or str(t.dtype).replace("torch.", "") != e["dtype"] \
                or hashlib.sha256(t.numpy().tobytes()).hexdigest() != e["sha256"]:
            bad.append(e["name"])
    assert not bad, bad[:10]


def test_decode_planners():
    """host-side planning of the cached generation step (no GPU): kv-split choices of the two decode attention paths, the
    unit decomposition of the persistent step kernel, and the packed weight image it streams"""
    from vae_gslm_b200.decode_step import attention_splits, pack_units, split_factor
    from vae_gslm_b200.ops import decode_splits
    # vg_attn_decode: splits cost more than the imbalance they remove at every cache length of the recipe
    assert all(decode_splits(b * 16, 683) == 1 for b in (1, 8, 64, 256))
    assert decode_splits(16, 4096) > 1 and decode_splits(4096, 4096) == 1
    # the step kernel's warp-per-item attention: the fitted cost model reproduces the measured optima
    n_warps = 148 * 7
    assert attention_splits(16 * 16, n_warps) == 4 and attention_splits(24 * 16, n_warps) == 2
    assert attention_splits(32 * 16, n_warps) == 2 and attention_splits(256 * 16, n_warps) == 1
    # units: R output features x K / S inputs, about 128 per phase, k-blocks per unit a power of two
    for (N, K) in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096), (2048, 1024), (1024, 64)):
        for rmax in (128, 256):
            R, S = split_factor(N, K, rmax=rmax)
            nkb = K // 64 // S
            assert R % 16 == 0 and R <= rmax and nkb >= 1 and nkb & (nkb - 1) == 0
            assert -(-N // R) * S <= 148
    # the packed image is the SWIZZLE_128B K-major layout of each [R x 64] slab: row r of a slab keeps its 16-byte group g
    # at position g ^ (r & 7)
    W = torch.arange(32 * 128, dtype=torch.float32).view(32, 128).to(torch.bfloat16)
    p = pack_units(W, 16, 2)                                  # 2 slabs x 2 k-slices x (1 k-block * 16 rows * 64)
    assert p.shape == (2, 2, 16 * 64)
    slab = p[1, 1].view(16, 8, 8)                             # rows 16..31, k 64..127
    for r in (0, 3, 9):
        for g in (0, 5):
            assert torch.equal(slab[r, g ^ (r & 7)], W[16 + r, 64 + g * 8: 64 + g * 8 + 8])
