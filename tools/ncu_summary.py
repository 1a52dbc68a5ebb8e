"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum) per kernel: launches, total ms, share."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>9s} {'total_ms':>12s} {'share':>8s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:60]:60s} {cnt[k]:9d} {tot[k]:12.3f} {100 * tot[k] / total:7.2f}%")
print(f"{'TOTAL':60s} {sum(cnt.values()):9d} {total:12.3f}")
