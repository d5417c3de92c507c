"""Instruction mix of a kernel from an `ncu --set full --import-source on` report.

    ncu -i REPORT.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_instmix.py src.csv [kernel-substring]

Prints, per SASS opcode: executed warp instructions, share, stall-sample share, shared-memory wavefronts."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(open(path)))
    kernels = []   # (name, header, rows)
    cur = None
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Kernel Name":
            cur = [r[1], rows[i + 1], []]
            kernels.append(cur)
            i += 2
            continue
        if cur is not None and r:
            cur[2].append(r)
        i += 1
    for name, hdr, body in kernels:
        if want not in name:
            continue
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        iw = hdr.index("L1 Wavefronts Shared")
        ops, samp, wf = collections.Counter(), collections.Counter(), collections.Counter()
        tot = ts = 0
        for r in body:
            if len(r) <= max(ia, ie, isamp, iw):
                continue
            toks = r[ia].strip().split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.rstrip(";")
            base = op.split(".")[0]
            if base in ("LDS", "STS", "LDG", "STG", "FADD2", "FMUL2", "FFMA2", "SYNCS", "UTCBAR", "LDTM", "STTM", "MUFU"):
                base = ".".join(op.split(".")[:2]) if base in ("LDS", "STS", "MUFU", "SYNCS") else base
            try:
                n = int(r[ie] or 0)
                sm = int(r[isamp] or 0)
            except ValueError:
                continue
            ops[base] += n
            samp[base] += sm
            tot += n
            ts += sm
            try:
                wf[base] += int(r[iw] or 0)
            except ValueError:
                pass
        print(f"# {name}\n# total warp instructions {tot}, stall samples {ts}")
        print(f"{'opcode':16s} {'warp-inst':>12s} {'share':>7s} {'samples':>8s} {'smem wavefronts':>16s}")
        for k, v in ops.most_common(60):
            print(f"{k:16s} {v:12d} {100 * v / max(tot, 1):6.2f}% {100 * samp[k] / max(ts, 1):7.2f}% {wf[k]:16d}")


if __name__ == "__main__":
    main()
