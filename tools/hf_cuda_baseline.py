"""The existing GPU implementation on the same box: HF's own `device="cuda"` branch of the extractor
(cuFFT + cuBLAS + ATen elementwise, TF-FE:140-163).  The reference never uses it; it is the second baseline of
SURVEY.md section 8d.  Prints audio-seconds/second for a device-resident batch (the H2D of the PCM and the D2H of
the features that `_torch_extract_fbank_features` performs are excluded by calling its body on device tensors)."""
import sys
import time

import numpy as np
import torch
from transformers import WhisperFeatureExtractor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_mels = int(sys.argv[2]) if len(sys.argv) > 2 else 80
fe = WhisperFeatureExtractor(feature_size=n_mels)
dev = torch.device("cuda")
wave = (0.1 * torch.randn(B, 480000, device=dev))
window = torch.hann_window(400, device=dev)
mel = torch.from_numpy(fe.mel_filters).to(dev, torch.float32)


def step():
    stft = torch.stft(wave, 400, 160, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    spec = mel.T @ mag
    log_spec = torch.clamp(spec, min=1e-10).log10()
    mx = log_spec.max(dim=2, keepdim=True)[0].max(dim=1, keepdim=True)[0]
    log_spec = torch.maximum(log_spec, mx - 8.0)
    return (log_spec + 4.0) / 4.0


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for _ in range(K):
    out = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
# also the full public call with device="cuda" (includes H2D of PCM and D2H of features, as the class does it)
host = wave.cpu().numpy()
t0 = time.perf_counter()
r = fe(host, sampling_rate=16000, device="cuda").input_features
t1 = time.perf_counter()
print(f"HF cuda body: B={B} n_mels={n_mels} {ms:.3f} ms/step -> {30.0 * B / (ms * 1e-3):.0f} audio-s/s device-resident; "
      f"public call with device='cuda' (H2D + D2H inside): {30.0 * B / (t1 - t0):.0f} audio-s/s")
