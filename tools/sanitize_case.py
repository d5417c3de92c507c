"""Smallest case that touches every kernel path (edge tiles, ragged lengths, int16, empty clip, both banks,
table-driven bank) for `compute-sanitizer --tool memcheck`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor  # noqa: E402

rng = np.random.default_rng(0)
clips = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in (480000, 0, 12345, 100001)]
for m in (80, 128, 64):
    fe = B200WhisperFeatureExtractor(feature_size=m)
    a = fe(clips, sampling_rate=16000, return_tensors="np").input_features
    q = [np.round(c * 32767).astype(np.int16) for c in clips]
    b = fe.extract_host(q).cpu().numpy()
    print(m, a.shape, float(np.abs(a).max()), float(np.abs(a - b).max()))
    fe.close()
print("done")
