#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("at::native::", "")
    k = re.sub(r"<unnamed>::", "", k)[:84]
    agg[k][0] += 1
    agg[k][1] += float(row["Metric Value"].replace(",", ""))
    n += 1
tot = sum(v[1] for v in agg.values())
mine = sum(v[1] for k, v in agg.items() if k.startswith("vx::"))
print("%d launches; GPU time %.2f ms; vx:: kernels %.2f ms (%.1f%%) in %d launches" %
      (n, tot / 1e6, mine / 1e6, 100 * mine / tot, sum(v[0] for k, v in agg.items() if k.startswith("vx::"))))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-86s n=%4d %8.1f us %5.1f%%  avg %6.1f" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / 1e3 / v[0]))
