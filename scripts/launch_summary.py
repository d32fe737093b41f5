"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step.
    python scripts/launch_summary.py gpurun_out/train_launches.csv [steps_in_capture]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
ks = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[start:] if len(r) > vi]
step = ks[len(ks) - len(ks) // steps:]
tot = sum(v for _, v, _ in step)
print("%d launches in the capture, %d in the last step, %.1f us" % (len(ks), len(step), tot / 1000))
agg = {}
for k, v, _ in step:
    k = re.sub(r"\(.*", "", k).replace("aae::", "").replace("<unnamed>::", "").replace("void ", "")[:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%3d  %9.1f us  %5.1f%%" % (k, c, v / 1000, 100 * v / tot))
if "-v" in sys.argv:
    for k, v, g in step:
        k = re.sub(r"\(.*", "", k).replace("aae::", "").replace("<unnamed>::", "").replace("void ", "")[:44]
        if v > 30000:
            print("   %-46s %-18s %8.1f us" % (k, g, v / 1000))
