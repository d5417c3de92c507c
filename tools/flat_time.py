"""Launch time against the number of clips handed to the flat kernel (set_flat_clips; 'auto' = the library's own split).
    python tools/flat_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor  # noqa: E402

for M in (80, 128):
    fe = B200WhisperFeatureExtractor(feature_size=M)
    g = torch.Generator(device="cuda").manual_seed(0)
    for B in (256, 1024):
        pcm = 0.1 * torch.randn(B, 480000, device="cuda", generator=g)
        out = torch.empty(B, M, 3000, device="cuda")
        for nflat in (0, 16, 32, 48, 64, 80, 96, "auto"):
            fe.set_flat_clips(-1 if nflat == "auto" else nflat)
            for _ in range(3):
                fe.extract_device(pcm, out=out)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fe.extract_device(pcm, out=out)
            e1.record()
            torch.cuda.synchronize()
            print(f"M={M} B={B:5d} flat={nflat}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per launch")
        del pcm, out
    fe.close()
