"""print the last N rows (kernel, grid, ns) of an ncu `--metrics gpu__time_duration.sum --csv` launch list"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
tot = 0.0
for r in rows[-n:]:
    v = float(r[vi].replace(",", ""))
    tot += v
    print(f"{r[ki][:64]:64s} {r[gi]:14s} {v / 1e3:9.1f} us")
print(f"sum of the last {n}: {tot / 1e3:.1f} us")
