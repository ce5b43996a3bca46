"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
  python tools/launch_summary.py gpurun_out/x_ncu_launches.csv [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    n = r[ki].split("(")[0]
    agg[n][0] += 1
    agg[n][1] += float(r[vi].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%4d %9.2f ms %5.1f%%" % (k[:72], v[0], v[1], 100 * v[1] / tot))
print("launches %d, total %.2f ms" % (len(data), tot))
