"""Launch time of dense batches with and without the flat kernel (set_flat_clips 0 / -1), per round of co-resident clusters.
    python tools/perf_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor  # noqa: E402

cases = [(80, 242), (80, 256), (128, 242), (128, 1024)]
if len(sys.argv) > 1:
    cases = [(int(sys.argv[1]), int(sys.argv[2]))]
for M, B in cases:
    fe = B200WhisperFeatureExtractor(feature_size=M)
    g = torch.Generator(device="cuda").manual_seed(0)
    pcm = 0.1 * torch.randn(B, 480000, device="cuda", generator=g)
    out = torch.empty(B, M, 3000, device="cuda")
    for nflat in (0, -1):
        fe.set_flat_clips(nflat)
        for _ in range(3):
            fe.extract_device(pcm, out=out)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        n0 = fe.launch_count
        e0.record()
        for _ in range(20):
            fe.extract_device(pcm, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"M={M} B={B:5d} flat={nflat:2d}: {us:8.1f} us per step, {us / (B / fe.max_clusters):6.2f} us per round of "
              f"{fe.max_clusters} clips, launches/step {(fe.launch_count - n0) / 20:.0f}", flush=True)
    del pcm, out
    fe.close()
