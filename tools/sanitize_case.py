"""Smallest meaningful case for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
Dense batch (static split incl. the flat kernel), ragged batch (clip queue, both kernels), host path, 16-bit store."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import logmel_oracle as O  # noqa: E402
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor  # noqa: E402

for m, dt in ((80, torch.float32), (128, torch.float16)):
    fe = B200WhisperFeatureExtractor(feature_size=m, feature_dtype=dt)
    g = torch.Generator(device="cuda").manual_seed(0)
    pcm = 0.1 * torch.randn(60, 480000, device="cuda", generator=g)
    lens = torch.tensor([0, 1, 5000, 5121, 100000, 479999, 480000, 33333, 160, 7777] * 6, dtype=torch.int32, device="cuda")
    a = fe.extract_device(pcm)                       # static: 44 + 16 clips
    fe.set_flat_clips(9)
    b = fe.extract_device(pcm, lengths=lens)         # queue, flat kernel takes part
    fe.set_flat_clips(-1)
    c = fe.extract_device(pcm, lengths=lens)
    clips = [O.synth_clip("speech", 48000, 1), O.synth_clip("noise", 480000, 2), O.synth_clip("zeros", 100, 3)]
    d = fe(clips, sampling_rate=16000).input_features
    torch.cuda.synchronize()
    assert torch.equal(b, c) and torch.isfinite(a.float()).all() and torch.isfinite(d.float()).all()
    ref = O.extract(clips, m, "f64")
    print(m, dt, "max-abs vs oracle", float(np.abs(d.float().cpu().numpy() - ref).max()))
    fe.close()
print("sanitize case done")
