"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    rows.append((name[:90], v * scale))
tot = sum(t for _, t in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, t in rows:
    agg[n][0] += 1
    agg[n][1] += t
print("%d launches, %.3f ms total (cold-cache, serialised under ncu)" % (len(rows), tot / 1e3))
print("%-92s %6s %10s %7s" % ("kernel", "count", "total us", "share"))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-92s %6d %10.1f %6.1f%%" % (n, c, t, 100 * t / tot))
