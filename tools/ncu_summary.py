"""Summarise an `ncu --set full` report into the numbers the design reasons about.

    python tools/ncu_summary.py REPORT.ncu-rep [--tag r02a_c2_80mel] [--workload "C2 ..."] [--clips 256]

Prints a text summary; with --tag also writes profiles/<tag>_summary.json, which bench.py reads for `roofline.traffic`
and the pipe utilisations.  The JSON records the sha of the kernel sources the profiled library was built from
(whisper_context_biasing_b200/build.py::kernel_sources_sha at the time this tool runs: run it BEFORE editing the kernels
again), so a stale summary is detected instead of printed."""
import argparse
import csv
import importlib.util
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag")
    ap.add_argument("--workload", default="")
    ap.add_argument("--clips", type=int, default=0)
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        k = {"name": name}
        print("kernel:", name[:90])
        for key in KEYS:
            if key in d and d[key] not in ("", "n/a"):
                print(f"  {key:85s} {d[key]} {u.get(key, '')}")
                try:
                    k[key] = to_bytes(d[key], u[key]) if key.startswith("dram__bytes") else float(d[key].replace(",", ""))
                except ValueError:
                    pass
        st = {h: float(v) for h, v in d.items()
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v not in ("", "n/a")}
        tot = sum(st.values()) or 1.0
        k["stalls_pct"] = {}
        for h, v in sorted(st.items(), key=lambda x: -x[1])[:9]:
            nm = h.replace("smsp__pcsamp_warps_issue_stalled_", "")
            print(f"  stall {nm:30s} {100 * v / tot:5.1f}%")
            k["stalls_pct"][nm] = round(100 * v / tot, 2)
        kernels.append(k)
    if args.tag:
        spec = importlib.util.spec_from_file_location("_wlm_build", os.path.join(ROOT, "whisper_context_biasing_b200", "build.py"))
        B = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(B)
        doc = {"tag": args.tag, "workload": args.workload, "clips": args.clips,
               "kernel_sources_sha": B.kernel_sources_sha(),
               "dram_bytes_per_step": sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in kernels),
               "kernels": kernels,
               "how": "ncu --set full --clock-control none --import-source on, one step of bench.py (cold cache, kernels serialised)"}
        path = os.path.join(ROOT, "profiles", args.tag + "_summary.json")
        json.dump(doc, open(path, "w"), indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
