"""Live reference arm (TEST / BASELINE INFRASTRUCTURE, not product).

The reference repo executes the hot path through the *third-party* Hugging Face
`WhisperFeatureExtractor` (see `oracle/logmel_oracle.py` header).  `transformers` is part of this
image both in the build container and on the GPU box, so the unmodified implementation can be run
directly -- this is what `bench.py --impl reference`, the `cpu_baseline` leg, the golden generator
and the parity tests use.  `/root/reference` itself is never read here (it does not exist on the
GPU box); the three call-site lines of the reference are restated instead and cited.
"""
from __future__ import annotations

import os
import time

import numpy as np


def make_hf_extractor(n_mels: int = 80):
    """`WhisperFeatureExtractor(feature_size=n_mels)`: the default ctor is the whisper-base/small
    config the reference loads with `from_pretrained("openai/whisper-base[.en]")`
    (REF/data_utils/data_collator.py:10, REF/scripts/train.py:96); 128 = large-v3.
    `from_pretrained` needs the network, so the object is built from its args."""
    from transformers import WhisperFeatureExtractor

    return WhisperFeatureExtractor(feature_size=n_mels)


def hf_features(clips, n_mels: int = 80, path: str = "default") -> np.ndarray:
    """Run the live extractor the way the reference does: ONE CLIP PER CALL
    (REF/data_utils/data_loader.py:171) and stack (REF/data_utils/data_collator.py:64-76).

    path "default": the class' own dispatch (torch fp32 STFT when torch is importable).
    path "numpy":   forced `_np_extract_fbank_features` (fp64 fallback).
    """
    fe = make_hf_extractor(n_mels)
    outs = []
    for x in clips:
        x = np.asarray(x, dtype=np.float32)
        if path == "default":
            outs.append(fe(x, sampling_rate=16000).input_features[0])
        else:
            padded = np.zeros((1, fe.n_samples), dtype=np.float32)
            xx = x[: fe.n_samples]
            padded[0, : xx.shape[0]] = xx
            outs.append(fe._np_extract_fbank_features(padded, "cpu")[0].astype(np.float32))
    return np.stack(outs, axis=0)


# ----------------------------------------------------------------------------------------------
# CPU baseline: the reference's Dataset.__getitem__ + collator feature stack under a DataLoader
# ----------------------------------------------------------------------------------------------
class _RefStyleDataset:
    """`__getitem__` restates REF/data_utils/data_loader.py:170-172 with `librosa.load` replaced
    by the seeded synthetic PCM (audio decode is not on the measured path)."""

    def __init__(self, clips, n_mels, path):
        self.clips = clips
        self.n_mels = n_mels
        self.path = path
        self.fe = None

    def __len__(self):
        return len(self.clips)

    def __getitem__(self, i):
        import torch

        if self.fe is None:
            torch.set_num_threads(1)
            self.fe = make_hf_extractor(self.n_mels)
        audio = self.clips[i]
        if self.path == "default":
            processed = self.fe(audio, sampling_rate=16000).input_features      # :171
        else:
            padded = np.zeros((1, self.fe.n_samples), dtype=np.float32)
            xx = audio[: self.fe.n_samples]
            padded[0, : xx.shape[0]] = xx
            processed = self.fe._np_extract_fbank_features(padded, "cpu").astype(np.float32)
        return {"input_features": torch.tensor(processed[0])}                   # :172


class _RefStyleCollate:
    """Feature half of REF/data_utils/data_collator.py:64-76 (`feature_extractor.pad(...,
    padding="longest", return_tensors="pt")`)."""

    def __init__(self, n_mels):
        self.n_mels = n_mels
        self.fe = None

    def __call__(self, features):
        if self.fe is None:
            self.fe = make_hf_extractor(self.n_mels)
        inp = {"input_features": [f["input_features"] for f in features]}
        return self.fe.pad(inp, padding="longest", return_tensors="pt")


class _EpochBatches:
    """batch_sampler living in the main process: lets the warm-up epoch be short while the timed
    epoch covers every clip, with the same persistent workers."""

    def __init__(self, n, batch_size):
        self.n, self.bs, self.limit = n, batch_size, None

    def __iter__(self):
        n = self.n if self.limit is None else min(self.n, self.limit)
        for i in range(0, n, self.bs):
            yield list(range(i, min(i + self.bs, n)))

    def __len__(self):
        n = self.n if self.limit is None else min(self.n, self.limit)
        return (n + self.bs - 1) // self.bs


def time_reference_dataloader(clips, n_mels: int, batch_size: int = 16, num_workers: int | None = None,
                              path: str = "default", epochs: int = 1, warmup_epochs: int = 1):
    """Returns dict(audio_s_per_s, seconds, epoch_seconds, clips, cores).  Audio-seconds are NOMINAL 30 s
    windows per clip (the extractor always pads/trims to 30 s).

    ONE DataLoader with persistent workers for the whole measurement: `warmup_epochs` untimed short epochs (two batches
    per worker) pay for worker start-up and extractor construction, then `epochs` full epochs are timed, each from the
    creation of its iterator, so prefetching cannot hide work from the clock.  `clips` in the result is per epoch."""
    from torch.utils.data import DataLoader

    if num_workers is None:
        num_workers = len(os.sched_getaffinity(0))
    sampler = _EpochBatches(len(clips), batch_size)
    dl = DataLoader(_RefStyleDataset(clips, n_mels, path), batch_sampler=sampler, num_workers=num_workers,
                    collate_fn=_RefStyleCollate(n_mels), persistent_workers=num_workers > 0,
                    prefetch_factor=2 if num_workers > 0 else None)
    sampler.limit = 2 * batch_size * max(1, num_workers)
    for _ in range(max(1, warmup_epochs)):
        for _ in dl:
            pass
    sampler.limit = None
    secs = []
    n_timed = 0
    for _ in range(max(1, epochs)):
        t0 = time.perf_counter()
        n_timed = 0
        for b in dl:
            n_timed += b["input_features"].shape[0]
        secs.append(time.perf_counter() - t0)
    del dl
    total = sum(secs)
    return {"audio_s_per_s": 30.0 * n_timed * len(secs) / total if total > 0 else 0.0, "seconds": total,
            "epoch_seconds": secs, "clips": n_timed, "cores": max(1, num_workers)}
