"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel total time and share.
    python tools/launch_summary.py gpurun_out/launches.csv [--skip N] [--last-pass K]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", "")) / 1e3) for r in csv.DictReader(lines)]
rows = rows[skip:]
if "--half" in sys.argv:
    rows = rows[len(rows) // 2:]
agg = OrderedDict()
for n, g, t in rows:
    key = n.split("(")[0].replace("dsir::", "").replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':50s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {c:8d} {t:10.1f} {100 * t / tot:6.1f}%")
print(f"{'TOTAL':50s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")
