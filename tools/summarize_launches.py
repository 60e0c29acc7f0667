"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    tot[name][0] += 1
    tot[name][1] += val * scale
total = sum(v[1] for v in tot.values())
print("| kernel | launches | total us | share | avg us |")
print("|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.0f | %.1f%% | %.1f |" % (name[:90], n, t, 100 * t / total, t / n))
print("| **total** | %d | %.0f | 100%% | |" % (sum(v[0] for v in tot.values()), total))
