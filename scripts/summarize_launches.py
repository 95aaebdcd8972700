"""prints the per-kernel breakdown of the LAST multiplication in an ncu launch-list CSV (gpurun_out/launches.csv)"""
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, gi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Name")
items = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[start:] if r[mi] == "gpu__time_duration.sum"]
n = len(items) // steps
step = items[-n:]
tot = sum(v for _, v, _ in step)
print(f"# {n} launches, sum of kernel durations {tot / 1e3:.1f} us")
for name, v, g in step:
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-44:]
    print(f"{v / 1e3:9.1f} us {100 * v / tot:5.1f}%  grid={g:>16}  {short}")
