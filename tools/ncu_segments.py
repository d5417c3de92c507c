"""Per-code-segment instruction / stall-sample shares from an .ncu-rep source page (SASS).
Segments are cut at barriers; pass marker regexes to name regions."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
ntiles = float(sys.argv[2]) if len(sys.argv) > 2 else 12032.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = rows[2:]
iS, iA, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[iA]) for r in data)
totE = sum(int(r[iE]) for r in data)
print("total samples", tot, "warp-instr", totE, "per tile", totE / ntiles)
start = 0
acc = accE = 0
for i, r in enumerate(data):
    acc += int(r[iA]); accE += int(r[iE])
    if re.search(r"BAR\.SYNC|UCGABAR_WAIT|SYNCS\.PHASECHK|STTM|LDTM", r[iS]) or i == len(data) - 1:
        if accE / totE > 0.002 or acc / tot > 0.002:
            print(f"{start:5d}-{i:5d} samples {100 * acc / tot:5.1f}%  instr {100 * accE / totE:5.1f}% ({accE / ntiles:7.0f}/tile)  ends: {r[iS].strip()[:60]}")
        start = i + 1; acc = accE = 0
print("top stalled:")
top = sorted(range(len(data)), key=lambda i: -int(data[i][iA]))[:18]
for i in sorted(top):
    print(f"  {i:5d} {data[i][iA]:>6s} {data[i][iE]:>9s} {data[i][iS].strip()[:100]}")
