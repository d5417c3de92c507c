"""Summarise an .ncu-rep (raw page) into the handful of numbers the design reasons about."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get("Kernel Name", "?")[:70])
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
    for k in keys:
        if k in d:
            print(f"  {k:85s} {d[k]}")
    st = {h: float(v) for h, v in d.items() if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v not in ("", "n/a")}
    tot = sum(st.values()) or 1.0
    for h, v in sorted(st.items(), key=lambda x: -x[1])[:9]:
        print(f"  stall {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):30s} {100 * v / tot:5.1f}%")
