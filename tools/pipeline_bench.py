"""can the input pipeline feed the training step?  A synthetic on-disk corpus in the recipe's format (token
metadata lines + pre-computed .npy mels of LibriLight-like lengths, 80 mel bins) is read through TokenMelDataset (crop to
token_segment_size = 640, utterance crop 2-4 s, post_pad) by `num_workers` DataLoader workers (the reference uses 6:
configs/train/speech/vae-gslm.yaml) and assembled into the pinned, channel-interleaved batches TrainStep.load() consumes.
Reports mel frames per second of the HOST pipeline, to put next to the GPU step rate (BENCH: ~4.0e5 frames/s at 8 x 640).
usage: python tools/pipeline_bench.py [workers] [n_utterances] [batches]"""
import os
import sys
import tempfile
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vae_gslm_b200.data.dataset import BatchAssembler, TokenMelDataset, train_batches
from vae_gslm_b200.hparams.hp import Hparams

workers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_utt = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_batches = int(sys.argv[3]) if len(sys.argv) > 3 else 120
N_MELS, B, T = 80, 8, 640
tmp = tempfile.mkdtemp(prefix="vg_pipe_")
rng = np.random.default_rng(3)
wavdir, meldir = os.path.join(tmp, "wav"), os.path.join(tmp, "mels")
lines = []
for i in range(n_utt):
    sub = f"spk{i % 16}"
    os.makedirs(os.path.join(wavdir, sub), exist_ok=True)
    os.makedirs(os.path.join(meldir, sub), exist_ok=True)
    frames = int(rng.integers(700, 3000))                      # 14 - 60 s utterances at 50 Hz
    with open(os.path.join(wavdir, sub, f"utt{i}.flac"), "wb") as f:
        f.truncate(int(18500 * frames / 50))                    # sparse file: only its SIZE is read (length filter)
    np.save(os.path.join(meldir, sub, f"utt{i}.npy"), rng.normal(size=(frames, N_MELS)).astype(np.float32))
    lines.append(f"{sub}/utt{i}.flac|" + " ".join(str(int(t)) for t in rng.integers(0, 200, frames)))
meta = os.path.join(tmp, "tokens.txt")
open(meta, "w").write("\n".join(lines) + "\n")
cfg = {"path": meta, "wavdir": wavdir, "preprocess_mels": meldir, "preprocess_mels_recursive_dir": True,
       "sample_rate": 16000, "with_text": False, "with_tokens": True, "min_audio_length": 1.0, "bits_per_second": 18500,
       "token_segment_size": T, "random_crop_mel_utt": {"min_seg_sec": 2.0, "max_seg_sec": 4.0},
       "post_pad": {"tokens": {"num_tokens": T}, "mel": {"length": T / 50.0}, "cropped_mel_utt": {"length": 4.0}}}
mel = {"sample_rate": 16000, "n_fft": 1024, "hop_length": 320, "n_mels": N_MELS, "power": 1}
ds = TokenMelDataset(Hparams.from_dict(cfg), Hparams.from_dict(mel), Hparams.from_dict({"deduplicate": False, "sample_rate": 50}),
                     Hparams.from_dict({"mean": -1.5, "std": 2.0}))
asm = BatchAssembler(B, T, 200, N_MELS, depth=3, pin=torch.cuda.is_available())
for w in sorted({0, workers}):
    done, t0, frames = 0, None, 0
    epoch = 0
    while done < n_batches + 10:
        for batch in train_batches(ds, asm, B, shuffle=True, num_workers=w, seed=1, epoch=epoch):
            if done == 10:
                t0 = time.perf_counter()                        # first batches: worker start-up, page cache
            if done >= 10:
                frames += int(batch["mask"].sum())
            done += 1
            if done >= n_batches + 10:
                break
        epoch += 1
    dt = time.perf_counter() - t0
    print(f"workers {w}: {n_batches} batches of {B} x {T} frames in {dt:.2f} s = {frames / dt:,.0f} valid mel frames/s "
          f"({n_batches / dt:.1f} batches/s; host cpus {os.cpu_count()})", flush=True)
