"""Run the C2 workload once with a -DWLM_TRACE build and dump CTA 0's phase timeline (gpurun_out/trace.npy).
    WLM_LIBRARY_PATH=build/libwlm_trace.so python tools/trace_run.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor, _native as N  # noqa: E402

fe = B200WhisperFeatureExtractor(feature_size=80)
g = torch.Generator(device="cuda").manual_seed(0)
pcm = 0.1 * torch.randn(256, 480000, device="cuda", generator=g)
for _ in range(3):
    out = fe.extract_device(pcm)
torch.cuda.synchronize()
n = 16 * 48 * 12
buf = (ctypes.c_ulonglong * n)()
N.LIB.wlm_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
rc = N.LIB.wlm_debug_trace(buf, n)
tr = np.frombuffer(buf, dtype=np.uint64).reshape(16, 48, 12).astype(np.int64)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/trace.npy", tr)
n2 = 16 * 8 * 4
buf2 = (ctypes.c_ulonglong * n2)()
N.LIB.wlm_debug_trace_clip.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
N.LIB.wlm_debug_trace_clip(buf2, n2)
np.save("gpurun_out/trace_clip.npy", np.frombuffer(buf2, dtype=np.uint64).reshape(16, 8, 4).astype(np.int64))
print("rc", rc, "span cycles", tr[tr > 0].max() - tr[tr > 0].min())
