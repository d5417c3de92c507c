"""Where the time of the reference's call shape goes (one pageable 30 s clip per call): the whole call, the pageable H2D copy
alone, a copy through a pinned buffer, the kernel alone.   python tools/c1_probe.py   (needs a B200)
Measured (16-core box): call 184 us, pageable H2D 106 us, memcpy to pinned 111 us (+ H2D: 163 us), kernel 31 us -- the driver's
pageable path is as fast as a copy through our own pinned ring would be, which is why small calls use it directly."""
import time, numpy as np, torch, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor
fe = B200WhisperFeatureExtractor(feature_size=80)
rng = np.random.default_rng(0)
clips = [(0.1 * rng.standard_normal(480000)).astype(np.float32) for _ in range(16)]
pin = torch.empty(480000, dtype=torch.float32).pin_memory()
dev = torch.empty(480000, device='cuda')
def t(f, n=5):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e6 / 16
print("fe(x) per clip            %.0f us" % t(lambda: [fe(c, sampling_rate=16000).input_features for c in clips]))
print("pageable H2D only         %.0f us" % t(lambda: [dev.copy_(torch.from_numpy(c), non_blocking=True) for c in clips]))
print("memcpy to pinned + H2D    %.0f us" % t(lambda: [(pin.numpy().__setitem__(slice(None), c), dev.copy_(pin, non_blocking=True), torch.cuda.synchronize()) for c in clips]))
print("memcpy to pinned only     %.0f us" % t(lambda: [pin.numpy().__setitem__(slice(None), c) for c in clips]))
d = torch.from_numpy(np.stack(clips)).cuda()
print("kernel only (device clip) %.0f us" % t(lambda: [fe.extract_device(d[i:i+1]) for i in range(16)]))
