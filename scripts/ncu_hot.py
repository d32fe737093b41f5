"""Top stall sites of a kernel from an ncu report's SASS source page:  python scripts/ncu_hot.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci, si, xi = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > ci]
tot = sum(float(r[ci] or 0) for r in body)
print("total samples", tot)
for idx, r in sorted(enumerate(body), key=lambda t: -float(t[1][ci] or 0))[:n]:
    print("%5.1f%%  #%-5d exec=%-8s %s" % (100 * float(r[ci] or 0) / tot, idx, r[xi], r[si].strip()[:100]))
