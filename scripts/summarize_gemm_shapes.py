"""Join an ncu launch list of one UNet3D call with the GEMM/conv shape log of the same call (scripts/profile_unet.py with
AP_SHAPE_LOG=...): per (op, M, N, K, kernel template) launch count, total time, average time and algorithmic TFLOP/s."""
import csv
import json
import re
import sys
from collections import defaultdict

launches, shapes = sys.argv[1], json.load(open(sys.argv[2]))
rows = []
with open(launches) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(r["Metric Unit"], 1)
    rows.append((r["Kernel Name"], ns))
g = [(n, t) for n, t in rows if "gemm_kernel" in n]
assert len(g) == len(shapes), (len(g), len(shapes))
agg = defaultdict(lambda: [0, 0.0, 0.0])
for (n, t), (kind, M, N, K, res) in zip(g, shapes):
    tmpl = re.search(r"gemm_kernel<([^>]*)>", n).group(1).replace(" ", "")
    a = agg[(kind, M, N, K, tmpl)]
    a[0] += 1
    a[1] += t
    a[2] += 2.0 * M * N * K
tot = sum(a[1] for a in agg.values())
print(f"{len(g)} GEMM-family launches, {tot / 1e6:.3f} ms (serialised, cold-cache ncu times), "
      f"{sum(a[2] for a in agg.values()) / 1e12:.2f} TFLOP")
print("   ms     n    avg us   TFLOP/s  (op, M, N, K, <BN,EPI,CG>)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / 1e6:7.3f} x{a[0]:3d} {a[1] / a[0] / 1e3:8.1f} {a[2] / a[1] / 1e3:8.0f}  {k}")
