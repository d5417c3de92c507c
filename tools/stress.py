"""Race hunt: the same random ragged / dense batches many times, through the cluster kernel alone and with a random share on
the flat kernel; every run must be bit-identical to the first (a missing hand-over shows up as a flipped bit sooner or later).
    python tools/stress.py [rounds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(123)
bad = 0
for m in (80, 128):
    fe = B200WhisperFeatureExtractor(feature_size=m)
    g = torch.Generator(device="cuda").manual_seed(m)
    for r in range(rounds):
        B = int(rng.integers(1, 200))
        pcm = 0.1 * torch.randn(B, 480000, device="cuda", generator=g)
        ragged = bool(rng.integers(0, 2))
        lens = None
        if ragged:
            kind = rng.integers(0, 3, size=B)
            L = np.where(kind == 0, rng.integers(0, 480001, size=B), np.where(kind == 1, rng.integers(0, 20000, size=B), 480000))
            lens = torch.tensor(L, dtype=torch.int32, device="cuda")
        fe.set_flat_clips(0)
        ref = fe.extract_device(pcm, lengths=lens).clone()
        for rep in range(4):
            nflat = int(rng.integers(0, B)) if rep else 0
            fe.set_flat_clips(nflat)
            got = fe.extract_device(pcm, lengths=lens)
            if not torch.equal(got, ref):
                bad += 1
                d = (got - ref).abs()
                print(f"MISMATCH m={m} round={r} B={B} ragged={ragged} flat={nflat} max={d.max().item():.3e} "
                      f"clips={torch.nonzero(d.amax(dim=(1, 2)) > 0).flatten().tolist()[:8]}")
        assert torch.isfinite(ref).all()
    fe.close()
print("stress done, mismatches:", bad)
sys.exit(1 if bad else 0)
