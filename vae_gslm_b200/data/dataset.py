"""Input pipeline for VAE-GSLM training (reference ``data/dataset.py:20-104,250-444`` and
``utils/helpers.py:36-135``): HuBERT k-means token metadata + pre-computed ``.npy`` mel files → cropped, rescaled,
right-padded batches of ``TensorMask``.

On-disk formats (unchanged, so prepared corpora are drop-in):
  * metadata: one utterance per line, ``<audio path relative to wavdir>|<space separated token ids>`` — the path is
    everything before the first ``|``, the ids are everything after the LAST ``|`` (dataset.py:54-61,84);
  * mels: ``<preprocess_mels>/<stem>.npy`` (or the audio's sub-directory under ``preprocess_mels`` with
    ``preprocess_mels_recursive_dir``), float array ``[frames, n_mels]`` at ``sample_rate / hop_length`` frames/s.
Only the pre-computed-mel mode of the recipe is built (``preprocess_mels`` set, ``with_text: false``,
``with_tokens: true``): the torchaudio front end that recomputes mels from audio is outside the hot path.

What differs from the reference: the same per-item semantics and the SAME random draws in the same order
(``np.random.rand`` for the utterance-crop length, ``torch.randint`` for the two crop offsets), so a seeded run yields
identical batches; but a batch is assembled directly into ONE channel-interleaved model input ``[B, T, 1 + n_mels]``
(token id as float ⊕ mel — what ``trainers/speech/lvtr.py:117-118`` builds on the GPU every step) inside reusable
PINNED host buffers (``BatchAssembler``), ready for an asynchronous H2D copy into the training step's static buffers
(``TrainStep.load``).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Any, Dict, Iterable, List, Mapping, Optional, Tuple

import numpy as np
import torch

from ..hparams.hp import Hparams
from ..utils.tensormask import TensorMask


# ------------------------------------------------------------------------------------------ helpers.py:36-135
def random_crop_1d(signal: torch.Tensor, sample_rate: float, min_crop_length_sec: float):
    """crop ``int(sec·rate)`` frames at a ``torch.randint`` offset; shorter signals pass through (helpers.py:36-53)."""
    n = int(min_crop_length_sec * sample_rate)
    if n >= len(signal):
        return signal, 0, len(signal)
    start = int(torch.randint(low=0, high=len(signal) - n + 1, size=()))
    return signal[start:start + n], start, start + n


def pad_1d(signal: torch.Tensor, length: int) -> torch.Tensor:
    """right-pad dimension 0 with zeros up to ``length`` (helpers.py:56-70)."""
    if len(signal) >= length:
        return signal
    out = signal.new_zeros((length,) + tuple(signal.shape[1:]))
    out[:len(signal)] = signal
    return out


def pad_to_max_length(batch: Iterable[Mapping[str, Any]], max_lengths: Optional[Mapping[str, int]] = None) -> Dict[str, Any]:
    """collate: every ≥1-D tensor field becomes a right-padded ``TensorMask`` (truncated to ``max_lengths[key]`` when
    given, else padded to the batch maximum); other fields are gathered into lists (helpers.py:80-135)."""
    batch = list(batch)
    max_lengths = dict(max_lengths or {})
    target: Dict[str, int] = {}
    for el in batch:
        for k, v in el.items():
            if isinstance(v, torch.Tensor) and v.dim() >= 1:
                target[k] = max_lengths[k] if k in max_lengths else max(target.get(k, 0), len(v))
    out: Dict[str, Any] = {}
    for k, n in target.items():
        vals, masks = [], []
        for el in batch:
            v = el[k][:n]
            masks.append(torch.arange(n) < len(v))
            vals.append(pad_1d(v, n))
        out[k] = TensorMask(torch.stack(vals), torch.stack(masks))
    for el in batch:
        for k, v in el.items():
            if not (isinstance(v, torch.Tensor) and v.dim() >= 1):
                out.setdefault(k, []).append(v)
    return out


# ------------------------------------------------------------------------------------------ dataset.py:20-104
def load_token_metadata(metadata: str, wavdir: str = "", min_audio_length: Optional[float] = None,
                        max_audio_length: Optional[float] = None, bits_per_second: Optional[int] = None,
                        max_token_length: int = 1000000, min_token_length: int = 0) -> Tuple[List[str], List[torch.Tensor]]:
    """parse the token metadata file; the audio-length filter estimates seconds from the FILE SIZE of the audio
    (``getsize / bits_per_second``), exactly like the reference (dataset.py:63-76)."""
    if min_audio_length is not None or max_audio_length is not None:
        assert bits_per_second is not None, "min/max_audio_length need bits_per_second"
    files: List[str] = []
    tokens: List[torch.Tensor] = []
    with open(metadata, "r", errors="ignore") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            parts = line.split("|", 1)
            if bits_per_second is not None:
                seconds = os.path.getsize(os.path.join(wavdir, parts[0])) / float(bits_per_second)
                if min_audio_length is not None and seconds < min_audio_length:
                    continue
                if max_audio_length is not None and seconds > max_audio_length:
                    continue
            ids = np.array(parts[-1].rsplit("|", 1)[-1].split(), dtype=np.int16)
            if len(ids) > max_token_length or len(ids) < min_token_length:
                continue
            files.append(parts[0])
            tokens.append(torch.from_numpy(ids))
    return files, tokens


class TokenMelDataset(torch.utils.data.Dataset):
    """``DiscreteTokenDataset`` of the reference (dataset.py:371-444 over :250-334) in pre-computed-mel mode."""

    def __init__(self, hp: Hparams, hp_mel: Hparams, hp_hubert: Hparams, hp_rescale: Optional[Hparams] = None) -> None:
        assert hp.get("with_tokens", False) and not hp.get("with_text", False)
        assert hp.get("preprocess_mels", None), "only the pre-computed mel mode of the recipe is built"
        assert not hp.has("segment_size") and not hp.has("truncate")
        if hp_hubert.get("deduplicate", False):
            raise NotImplementedError("deduplicated token streams are not used by the VAE-GSLM recipe")
        self.hp, self.hp_rescale = hp, hp_rescale
        self.mel_rate = float(hp_mel.sample_rate) / float(hp_mel.hop_length)        # features.py:78-79
        self.token_rate = float(hp_hubert.sample_rate)
        paths = [hp.path] if isinstance(hp.path, str) else list(hp.path)
        wavdirs = [hp.wavdir] if isinstance(hp.wavdir, str) else list(hp.wavdir)
        bps = hp.get("bits_per_second", None)
        bps = bps if isinstance(bps, list) else [bps] * len(paths)
        self.audios: List[str] = []
        self.tokens: List[torch.Tensor] = []
        for path, wavdir, b in zip(paths, wavdirs, bps):
            files, toks = load_token_metadata(path, wavdir, hp.get("min_audio_length", None),
                                              hp.get("max_audio_length", None), b,
                                              hp.get("max_token_length", 1000000), hp.get("min_token_length", 0))
            self.audios += [os.path.join(wavdir, f) for f in files]
            self.tokens += toks
        self.wavdir = wavdirs[0]

    def __len__(self) -> int:
        return len(self.audios)

    def mel_path(self, i: int) -> str:
        p = Path(self.audios[i])
        if self.hp.get("preprocess_mels_recursive_dir", False):               # dataset.py:283-288
            rel = str((p.parents[0] / (p.stem + ".npy")).resolve())[len(str(Path(self.wavdir).resolve())) + 1:]
            return os.path.join(self.hp.preprocess_mels, rel)
        return os.path.join(self.hp.preprocess_mels, p.stem + ".npy")

    def __getitem__(self, i: int) -> Dict[str, torch.Tensor]:
        mel = torch.from_numpy(np.load(self.mel_path(i)).astype(np.float32))
        if self.hp_rescale is not None:
            mel = (mel - self.hp_rescale.mean) / self.hp_rescale.std              # dataset.py:321-322
        ret: Dict[str, torch.Tensor] = {"mel": mel}
        for key, name in (("random_crop_mel", "cropped_mel"), ("random_crop_mel_utt", "cropped_mel_utt")):
            if self.hp.has(key):                                                   # dataset.py:326-341
                sub = self.hp.get(key)
                seg = np.random.rand() * (sub.max_seg_sec - sub.min_seg_sec) + sub.min_seg_sec
                ret[name] = random_crop_1d(mel, self.mel_rate, seg)[0]
        tokens = self.tokens[i].long()
        if tokens.size(0) < ret["mel"].size(0):                                    # dataset.py:389-390
            ret["mel"] = ret["mel"][:tokens.size(0)]
        if self.hp.has("token_segment_size"):                                      # dataset.py:391-417
            n = self.hp.token_segment_size
            if n <= len(tokens):
                start = int(torch.randint(low=0, high=len(tokens) - n + 1, size=()))
                tokens = tokens[start:start + n]
                s = int(float(start) / self.token_rate * self.mel_rate)
                e = s + int(float(n) / self.token_rate * self.mel_rate)
                ret["mel"] = pad_1d(ret["mel"], e)[s:e]
        ret["tokens"] = tokens
        return ret

    def post_pad(self) -> Optional[Dict[str, int]]:
        """fixed collate lengths of ``hp.post_pad`` (dataset.py:343-368,432-444)."""
        if not self.hp.has("post_pad"):
            return None
        pp, out = self.hp.post_pad, {}
        for key in ("mel", "cropped_mel", "cropped_mel_utt"):
            if pp.has(key):
                out[key] = int(pp.get(key).length * self.mel_rate)
        if pp.has("tokens"):
            out["tokens"] = pp.tokens.num_tokens
        return out or None

    def collate(self, batch: Iterable[Mapping[str, Any]]) -> Dict[str, Any]:
        return pad_to_max_length(batch, self.post_pad())


class PinnedBatch(dict):
    """one pinned host slot of the BatchAssembler; ``event`` is set by ``TrainStep.load`` behind the asynchronous H2D
    copies that read it, and waited for before the slot is refilled"""
    event = None


class BatchAssembler:
    """collate straight into pinned, reusable host buffers in the layout the training step consumes:
    ``x [B,T,1+n_mels]`` (token id as float ⊕ mel, trainers/speech/lvtr.py:117-118), ``mask [B,T]``,
    ``utterance [B,Tu,n_mels]``, ``utt_mask [B,Tu]``.  ``depth`` buffers rotate so that the copy of batch i can be in
    flight while batch i+1 is assembled; a slot is not refilled before the copies that read it have executed (the
    consumer — ``TrainStep.load`` — leaves a CUDA event in ``slot.event``; the host may be several replays ahead)."""

    def __init__(self, batch_size: int, frames: int, utt_frames: int, n_mels: int, depth: int = 2,
                 pin: Optional[bool] = None) -> None:
        pin = torch.cuda.is_available() if pin is None else pin
        mk = (lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt).pin_memory() if pin else torch.zeros(*s, dtype=dt))
        self.slots = [PinnedBatch(x=mk(batch_size, frames, 1 + n_mels), mask=mk(batch_size, frames, dt=torch.bool),
                                  utterance=mk(batch_size, utt_frames, n_mels),
                                  utt_mask=mk(batch_size, utt_frames, dt=torch.bool))
                      for _ in range(depth)]
        self.next = 0

    def __call__(self, items, utt_key: str = "cropped_mel_utt") -> Dict[str, torch.Tensor]:
        """``items``: a list of dataset items, or a batch already laid out by ``fill`` in a loader worker (a dict of the
        four tensors): either way the result is the next pinned slot, filled."""
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if slot.event is not None:
            slot.event.synchronize()              # the H2D copies of the batch last assembled here have executed
            slot.event = None
        if isinstance(items, Mapping):
            for k, v in items.items():
                slot[k].copy_(v)
            return slot
        return self.fill(slot, items, utt_key)

    @staticmethod
    def fill(slot: Mapping[str, torch.Tensor], items: List[Mapping[str, torch.Tensor]], utt_key: str = "cropped_mel_utt"):
        B, T, _ = slot["x"].shape
        Tu = slot["utterance"].shape[1]
        assert len(items) <= B
        for t in slot.values():
            t.zero_()
        for b, it in enumerate(items):
            n = min(T, len(it["tokens"]), len(it["mel"]))
            slot["x"][b, :n, 0] = it["tokens"][:n].float()
            slot["x"][b, :n, 1:] = it["mel"][:n]
            # token and mel masks coincide in the recipe (both cropped to token_segment_size); keep the shorter
            slot["mask"][b, :n] = True
            u = it[utt_key][:Tu]
            slot["utterance"][b, :len(u)] = u
            slot["utt_mask"][b, :len(u)] = True
        return slot


class _WorkerCollate:
    """collate_fn that runs INSIDE a DataLoader worker: lays the items of a batch out in the final interleaved layout, so
    that four tensors per BATCH cross the process boundary instead of four per ITEM (measured: 6 workers handing lists of
    items to the parent reach 154 k frames/s, less than the 499 k of one process — tools/pipeline_bench.py)."""

    def __init__(self, shapes: Mapping[str, tuple], utt_key: str) -> None:
        self.shapes, self.utt_key = dict(shapes), utt_key

    def __call__(self, items):
        slot = {k: torch.zeros(shape, dtype=dt) for k, (shape, dt) in self.shapes.items()}
        return BatchAssembler.fill(slot, items, self.utt_key)


def train_batches(dataset: TokenMelDataset, assembler: BatchAssembler, batch_size: int, shuffle: bool = True,
                  num_workers: int = 0, seed: int = 0, drop_last: bool = True, rank: Optional[int] = None,
                  world_size: Optional[int] = None, epoch: int = 0, utt_key: str = "cropped_mel_utt"):
    """iterate pinned, assembled batches (keys match ``TrainStep.static``): items are read, cropped and laid out in the
    interleaved layout by ``num_workers`` DataLoader workers (in the parent when ``num_workers`` is 0) and land in the
    rotating pinned slots — ``step.load(batch)`` then issues the asynchronous H2D copies.  With ``rank`` / ``world_size`` the epoch is
    sharded by torch's ``DistributedSampler`` exactly as the reference's ``StandardSampler(distributed=True)`` does
    (data/sampler.py:9-24, training_lib/trainer.py:52-65): pass the epoch number so that every rank reshuffles alike."""
    # with workers the interleaved layout is produced in the worker (one set of four tensors per batch crosses the process
    # boundary); without, items are laid out directly into the pinned slot
    collate = list
    if num_workers > 0:
        collate = _WorkerCollate({k: (tuple(v.shape), v.dtype) for k, v in assembler.slots[0].items()}, utt_key)
    if rank is not None:
        assert world_size is not None
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world_size, rank=rank,
                                                                  shuffle=shuffle, seed=seed, drop_last=drop_last)
        sampler.set_epoch(epoch)
        loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, sampler=sampler, num_workers=num_workers,
                                             collate_fn=collate, drop_last=drop_last)
    else:
        g = torch.Generator().manual_seed(seed + epoch)
        loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                                             collate_fn=collate, drop_last=drop_last, generator=g)
    for items in loader:
        yield assembler(items, utt_key=utt_key)
