"""Debug helper (GPU box): fused kernel vs the oracle, error broken down by frame tile / mel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import logmel_oracle as O  # noqa: E402
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor  # noqa: E402

n_mels = int(sys.argv[1]) if len(sys.argv) > 1 else 80
fe = B200WhisperFeatureExtractor(feature_size=n_mels)
clips = [O.synth_clip("noise", 480000, 1), O.synth_clip("speech", 100000, 2), O.synth_clip("sine", 480000, 3),
         O.synth_clip("chirp", 30000, 4), O.synth_clip("zeros", 1000, 5)]
got = fe(clips, sampling_rate=16000, return_tensors="np").input_features
ref = O.extract(clips, n_mels, "f64")
for b in range(len(clips)):
    d = np.abs(got[b] - ref[b])
    print(f"clip {b}: max {d.max():.3e} at (mel,frame) {np.unravel_index(d.argmax(), d.shape)}  nan={np.isnan(got[b]).sum()}")
    per_tile = d.reshape(n_mels, -1)[:, :2944].reshape(n_mels, 46, 64).max(axis=(0, 2))
    print("   per tile:", np.array2string(per_tile[:12], precision=1, max_line_width=200))
    per_mel = d.max(axis=1)
    print("   per mel :", np.array2string(per_mel[:16], precision=1, max_line_width=200))
    fr = d[:, :128].max(axis=0)
    print("   frames 0..127:", np.array2string(fr, precision=0, max_line_width=250))
