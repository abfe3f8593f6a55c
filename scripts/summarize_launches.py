"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[1] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns, *_ in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}%  x{c:4d}  avg {t / c / 1e3:8.1f} us  {n[:110]}")
