"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name) if not name.startswith("gemm_tcgen05") and "mcan" not in r["Kernel Name"] else re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v * scale))
tot = sum(t for _, t in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in rows:
    agg[n][0] += 1
    agg[n][1] += t
print("total %.1f us over %d launches" % (tot, len(rows)))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.1f us %5.1f%%  x%-4d %s" % (t, 100 * t / tot, c, n[:110]))
