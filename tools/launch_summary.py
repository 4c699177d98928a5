"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    name = re.sub(r"\(.*", "", row[ki])
    name = re.sub(r"^void ", "", name)
    v = float(row[vi].replace(",", ""))
    unit = row[hdr.index("Metric Unit")]
    if unit == "ns":
        v /= 1e3
    elif unit == "ms":
        v *= 1e3
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot/1e3:.3f} ms over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:4d}  {k[:100]}")
