#!/usr/bin/env python
"""Sum gpu__time_duration per kernel name from an `ncu --csv` launch list; second half of the launches only when argv[2] == 'half'
(the target scripts run one warm-up call first)."""
import csv
import collections
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0], v))
if len(sys.argv) > 2 and sys.argv[2] == "half":
    rows = rows[len(rows) // 2:]
agg = collections.OrderedDict()
for k, v in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:12.1f} us {a[0]:5d} x  {100 * a[1] / tot:5.1f}%  {k[:90]}")
print(f"{tot:12.1f} us total, {len(rows)} launches")
