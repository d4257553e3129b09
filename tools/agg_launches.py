#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (total, count, average)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
iK, iV = H.index("Kernel Name"), H.index("Metric Value")
seq = [(r[iK], float(r[iV].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > iV]
print(len(seq), "launches,", round(sum(v for _, v in seq) / 1e6, 3), "ms total")
agg = collections.OrderedDict()
for k, v in seq:
    k = k.split("(")[0][:100]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t/1e6:10.3f} ms total {n:5d}x  avg {t/n/1e3:9.1f} us  {k}")
