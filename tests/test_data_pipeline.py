"""the input pipeline (vae_gslm_b200/data/dataset.py) against the REAL reference dataset classes
(data/dataset.py) on a synthetic on-disk corpus, with identical seeds → identical items and collated batches.
The reference only exists in the build container; without it the differential test is skipped and the format /
property tests still run."""
import os
import sys

import numpy as np
import pytest
import torch

from vae_gslm_b200.data.dataset import train_batches, BatchAssembler, TokenMelDataset, load_token_metadata, pad_to_max_length
from vae_gslm_b200.hparams.hp import Hparams

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from ref_shim import import_reference, reference_available  # noqa: E402

N_MELS = 8


def _corpus(tmp_path, n=7):
    rng = np.random.default_rng(3)
    wavdir, meldir = tmp_path / "wav", tmp_path / "mels"
    lines = []
    for i in range(n):
        sub = wavdir / f"spk{i % 2}"
        sub.mkdir(parents=True, exist_ok=True)
        frames = int(rng.integers(40, 400))
        ntok = frames - int(rng.integers(0, 3))
        (sub / f"utt{i}.flac").write_bytes(b"\0" * int(18500 * frames / 50))
        (meldir / f"spk{i % 2}").mkdir(parents=True, exist_ok=True)
        np.save(meldir / f"spk{i % 2}" / f"utt{i}.npy", rng.normal(size=(frames, N_MELS)).astype(np.float32))
        lines.append(f"spk{i % 2}/utt{i}.flac|" + " ".join(str(int(t)) for t in rng.integers(0, 200, ntok)))
    lines.insert(3, "")                                            # blank lines are skipped
    meta = tmp_path / "tokens.txt"
    meta.write_text("\n".join(lines) + "\n")
    cfg = {"path": str(meta), "wavdir": str(wavdir), "preprocess_mels": str(meldir), "preprocess_mels_recursive_dir": True,
           "sample_rate": 16000, "with_text": False, "with_tokens": True, "min_audio_length": 1.0, "bits_per_second": 18500,
           "token_segment_size": 64, "random_crop_mel_utt": {"min_seg_sec": 0.5, "max_seg_sec": 1.5},
           "post_pad": {"tokens": {"num_tokens": 64}, "mel": {"length": 1.28}}}
    mel = {"sample_rate": 16000, "n_fft": 1024, "hop_length": 320, "n_mels": N_MELS, "power": 1}
    hubert = {"deduplicate": False, "sample_rate": 50}
    rescale = {"mean": -1.5, "std": 2.0}
    return cfg, mel, hubert, rescale


def _hp(cls, d):
    import json
    return cls.from_json(json.dumps(d)) if hasattr(cls, "from_json") else cls.from_dict(d)


def test_metadata_filters_and_shapes(tmp_path):
    cfg, mel, hubert, rescale = _corpus(tmp_path)
    files, toks = load_token_metadata(cfg["path"], cfg["wavdir"], 1.0, None, 18500)
    assert 0 < len(files) <= 7 and all(t.dtype == torch.int16 for t in toks)
    ds = TokenMelDataset(Hparams.from_dict(cfg), Hparams.from_dict(mel), Hparams.from_dict(hubert), Hparams.from_dict(rescale))
    assert ds.mel_rate == 50.0 and ds.post_pad() == {"mel": 64, "tokens": 64}
    torch.manual_seed(0)
    np.random.seed(0)
    items = [ds[i] for i in range(len(ds))]
    for it in items:
        assert it["tokens"].dtype == torch.int64 and len(it["tokens"]) <= 64 and it["mel"].shape[1] == N_MELS
        assert 25 <= len(it["cropped_mel_utt"]) <= 75 or len(it["cropped_mel_utt"]) == len(it["mel"])
    batch = ds.collate(items)
    assert batch["tokens"].value.shape == (len(ds), 64) and batch["mel"].value.shape == (len(ds), 64, N_MELS)
    assert bool((batch["mel"].value[~batch["mel"].mask] == 0).all())
    slot = BatchAssembler(len(ds), 64, 75, N_MELS, pin=False)(items)
    assert torch.equal(slot["x"][..., 0], torch.where(batch["tokens"].mask, batch["tokens"].value, 0).float())
    assert torch.equal(slot["x"][..., 1:], batch["mel"].value) and torch.equal(slot["mask"], batch["tokens"].mask)
    u = pad_to_max_length([{"u": it["cropped_mel_utt"]} for it in items], {"u": 75})["u"]
    assert torch.equal(slot["utterance"], u.value) and torch.equal(slot["utt_mask"], u.mask)


@pytest.mark.skipif(not reference_available(), reason="the reference tree is only present in the build container")
def test_items_and_batches_match_reference(tmp_path):
    cfg, mel, hubert, rescale = _corpus(tmp_path)
    _, RefHp, _ = import_reference()
    from data.dataset import DiscreteTokenDataset            # the reference's (sys.path set by import_reference)
    ref = DiscreteTokenDataset(_hp(RefHp, cfg), _hp(RefHp, mel), _hp(RefHp, hubert), _hp(RefHp, rescale))
    ours = TokenMelDataset(Hparams.from_dict(cfg), Hparams.from_dict(mel), Hparams.from_dict(hubert), Hparams.from_dict(rescale))
    assert len(ref) == len(ours) and [os.path.normpath(a) for a in ref.audios] == [os.path.normpath(a) for a in ours.audios]
    got = []
    for ds in (ref, ours):
        torch.manual_seed(123)
        np.random.seed(123)
        items = [ds[i] for i in range(len(ds))]
        batch = ds.seqCollate(items) if ds is ref else ds.collate(items)
        got.append((items, batch))
    (ri, rb), (oi, ob) = got
    for a, b in zip(ri, oi):
        assert set(a) == set(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k
    assert set(rb) == set(ob)
    for k in rb:
        assert torch.equal(rb[k].value, ob[k].value) and torch.equal(rb[k].mask, ob[k].mask), k


def test_items_and_batches_match_golden(tmp_path):
    """same check against the committed output of the reference (tests/golden/make_data_golden.py)."""
    cfg, mel, hubert, rescale = _corpus(tmp_path)
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "data_pipeline.pt"), weights_only=False)
    ds = TokenMelDataset(Hparams.from_dict(cfg), Hparams.from_dict(mel), Hparams.from_dict(hubert), Hparams.from_dict(rescale))
    assert [os.path.relpath(a, cfg["wavdir"]) for a in ds.audios] == gold["audios"]
    torch.manual_seed(123)
    np.random.seed(123)
    items = [ds[i] for i in range(len(ds))]
    for a, b in zip(gold["items"], items):
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    batch = ds.collate(items)
    assert set(batch) == set(gold["batch"])
    for k, (value, mask) in gold["batch"].items():
        assert torch.equal(batch[k].value, value) and torch.equal(batch[k].mask, mask), k


def test_train_batches_rotate_slots(tmp_path):
    cfg, mel, hubert, rescale = _corpus(tmp_path)
    ds = TokenMelDataset(Hparams.from_dict(cfg), Hparams.from_dict(mel), Hparams.from_dict(hubert), Hparams.from_dict(rescale))
    asm = BatchAssembler(3, 64, 75, N_MELS, depth=2, pin=False)
    seen = list(train_batches(ds, asm, 3, shuffle=True, seed=1))
    assert len(seen) == len(ds) // 3 and seen[0]["x"] is not seen[1]["x"]
    for b in seen:
        assert set(b) == {"x", "mask", "utterance", "utt_mask"} and b["x"].shape == (3, 64, 1 + N_MELS)
        assert bool(b["mask"].any(1).all()) and bool((b["x"][~b["mask"]] == 0).all())


def test_train_batches_shard_by_rank(tmp_path):
    """two ranks see disjoint halves of the epoch (torch DistributedSampler, as the reference's StandardSampler) and
    every rank draws the same permutation for the same epoch."""
    cfg, mel, hubert, rescale = _corpus(tmp_path, n=8)
    cfg.pop("random_crop_mel_utt")                     # deterministic items: identify utterances by their content
    cfg["random_crop_mel"] = {"min_seg_sec": 100.0, "max_seg_sec": 100.0}      # crop longer than any file = identity
    ds = TokenMelDataset(Hparams.from_dict(cfg), Hparams.from_dict(mel), Hparams.from_dict(hubert), Hparams.from_dict(rescale))
    n = len(ds)
    seen = []
    for rank in range(2):
        torch.manual_seed(0)
        asm = BatchAssembler(1, 64, 512, N_MELS, pin=False)
        firsts = [float(b["utterance"][0, 0, 0]) for b in
                  train_batches(ds, asm, 1, shuffle=True, seed=5, rank=rank, world_size=2, epoch=3, utt_key="cropped_mel")]
        assert len(firsts) == n // 2
        seen.append(firsts)
    assert not set(seen[0]) & set(seen[1]) and len(set(seen[0]) | set(seen[1])) == 2 * (n // 2)
    want = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=2, rank=0, shuffle=True, seed=5, drop_last=True)
    want.set_epoch(3)
    first_of = {i: float(((torch.from_numpy(__import__("numpy").load(ds.mel_path(i))) - rescale["mean"]) / rescale["std"])[0, 0])
                for i in range(n)}
    assert seen[0] == [pytest.approx(first_of[i]) for i in want]
