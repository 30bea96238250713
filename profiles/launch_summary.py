"""Launch list of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` -> per-kernel totals and shares.
usage: python profiles/launch_summary.py gpurun_out/launches_r01.csv > profiles/launches_r01_summary.txt"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    rows.append((r["Kernel Name"], r.get("Grid Size", ""), ms))
agg = OrderedDict()
for name, grid, ms in rows:
    k = (name, grid)
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + ms)
total = sum(t for _, t in agg.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none on `python bench.py --steps 2 --warmup 3 --skip-cpu --skip-roundtrip` "
      "(cold-cache, serialised: compare shares)")
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[-68:]:>68} grid {grid:<16} launches {n:3d} total {t:9.3f} ms share {100 * t / total:5.1f}%  per launch {t / n:8.3f} ms")
