"""Print CTA 0's phase timeline (gpurun_out/trace.npy from tools/trace_run.py) as per-warp phase durations."""
import sys
import numpy as np
tr = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.npy")
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, tr - t0, -1)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (10, 14)
ev = ["s1_begin", "raw_ok", "loaded", "go1", "fft1_done", "ystored", "s2_begin", "yfull_ok", "go2", "pstored", "mel_begin", "mel_end"]
for w in (0, 1, 3, 7, 8, 9, 11):
    print("warp", w)
    for n in range(lo, hi):
        r = tr[w, n]
        d = lambda a, b: (r[b] - r[a]) if r[a] >= 0 and r[b] >= 0 else -1
        print(f"  tile {n}: t={r[0]:7d} rawwait {d(0,1):5d} load1 {d(1,2):5d} wait {d(2,3):5d} fft1 {d(3,4):5d} st1 {d(4,5):5d} | "
              f"s2@{r[6]:7d} yfullwait {d(6,7):5d} load2+wait {d(7,8):5d} fft2+st {d(8,9):5d} | mel@{r[10]:7d} {d(10,11):5d}")
per = [tr[w, 30, 0] - tr[w, 10, 0] for w in range(16)]
print("period per half-tile step (cycles):", [int(p / 20) for p in per])
