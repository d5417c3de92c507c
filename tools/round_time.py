"""Launch time against batch size: marginal time per round of 22 clips (one per co-resident cluster) and the fixed cost per
launch (prologue, first TMA, read-back of the last clip).  B = 22 / 44 keep the PCM inside the 126 MB L2: same round time as
the HBM-resident sizes, i.e. the kernel does not wait for HBM.
    python tools/round_time.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200 import B200WhisperFeatureExtractor
fe = B200WhisperFeatureExtractor(feature_size=80)
g = torch.Generator(device="cuda").manual_seed(0)
for B in (22, 44, 88, 176, 264, 528):
    pcm = 0.1 * torch.randn(B, 480000, device="cuda", generator=g)
    out = torch.empty(B, 80, 3000, device="cuda")
    for _ in range(5): fe.extract_device(pcm, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fe.extract_device(pcm, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"B={B:4d}: {ms*1e3:8.1f} us per launch, {ms*1e3/(B/22):6.2f} us per round of 22 clips")
