"""Times wlm_logmel of ANY build of libwlm.so through the bare C ABI (to compare library builds on the same GPU box in the
same call):   python tools/abi_probe.py path/to/libwlm.so [n_mels B]..."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_context_biasing_b200.feature_extraction import slaney_mel_filters  # noqa: E402

lib = C.CDLL(sys.argv[1])
vp = C.c_void_p
lib.wlm_plan_create.argtypes = [C.c_int, C.c_int, vp, C.POINTER(vp)]
lib.wlm_logmel.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int64, C.c_int, vp, vp, vp, C.c_size_t, vp]
lib.wlm_workspace_bytes.restype = C.c_size_t
lib.wlm_workspace_bytes.argtypes = [vp, C.c_int]
lib.wlm_last_error.restype = C.c_char_p
cases = [(80, 242), (80, 256), (128, 1024)]
ragged = "--c4" in sys.argv
argv = [x for x in sys.argv if x != "--c4"]
if len(argv) > 3:
    cases = [(int(argv[2]), int(argv[3]))]
if ragged:          # bench.py's C4: 4096 variable-length clips, ragged buffer
    cases = [(80, 4096)]
torch.cuda.init()
for M, B in cases:
    table = np.ascontiguousarray(slaney_mel_filters(M).astype(np.float32))
    plan = vp()
    assert lib.wlm_plan_create(0, M, table.ctypes.data, C.byref(plan)) == 0, lib.wlm_last_error()
    g = torch.Generator(device="cuda").manual_seed(0)
    offs_t = lens_t = None
    if ragged:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench

        lens = bench.c4_lengths(B)
        offs = np.zeros(B, dtype=np.int64)
        offs[1:] = np.cumsum((lens[:-1] + 7) // 8 * 8)
        total = int(offs[-1] + (lens[-1] + 7) // 8 * 8)
        pcm = 0.1 * torch.randn(total, device="cuda", generator=g)
        offs_t = torch.from_numpy(offs).cuda()
        lens_t = torch.from_numpy(lens.astype(np.int32)).cuda()
    else:
        pcm = 0.1 * torch.randn(B, 480000, device="cuda", generator=g)
    out = torch.empty(B, M, 3000, device="cuda")
    ws = torch.empty(max(256, lib.wlm_workspace_bytes(plan, B)), dtype=torch.uint8, device="cuda")
    st = vp(torch.cuda.current_stream().cuda_stream)

    def run():
        rc = lib.wlm_logmel(plan, vp(pcm.data_ptr()), 0, vp(offs_t.data_ptr()) if ragged else None,
                            vp(lens_t.data_ptr()) if ragged else None, 0 if ragged else 480000, B, vp(out.data_ptr()), None,
                            vp(ws.data_ptr()), ws.numel(), st)
        assert rc == 0, lib.wlm_last_error()

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
    # the same launches with a device synchronisation after each one: no launch can overlap its predecessor
    iso = 1e9
    for rep in range(3):
        tot = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        iso = min(iso, tot / 10 * 1e3)
    print(f"{os.path.basename(sys.argv[1])}: M={M} B={B:5d}: {best:8.1f} us per step back to back, {iso:8.1f} us isolated  "
          f"checksum {float(out.double().sum()):.6f}", flush=True)
    lib.wlm_plan_destroy(plan)
