"""Per-shape comparison of two MCAN_BENCH_DUMP files (bench.py gemm_roofline records)."""
import json
import sys
from collections import defaultdict


def load(path):
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for r in json.load(open(path)):
        key = (tuple(r["shape"][:5]), tuple(r["shape"][5]))
        agg[key][0] += 1
        agg[key][1] += r["us"]
        agg[key][2] += r["flops"]
    return agg


a, b = load(sys.argv[1]), load(sys.argv[2])
ta = tb = 0.0
rows = []
for k in sorted(set(a) | set(b), key=lambda k: -(a.get(k, [0, 0, 0])[1])):
    x, y = a.get(k, [0, 0.0, 0.0]), b.get(k, [0, 0.0, 0.0])
    ta += x[1]
    tb += y[1]
    rows.append("%-34s x%-3d %8.1f -> %8.1f us  (%+6.1f)  %6.0f -> %6.0f TF/s  %s" % (
        k[0], max(x[0], y[0]), x[1], y[1], y[1] - x[1], x[2] / max(x[1], 1e-9) / 1e6, y[2] / max(y[1], 1e-9) / 1e6,
        ",".join(s for s in k[1] if s not in ("seed",))))
print("\n".join(rows))
print("total %.1f -> %.1f us" % (ta, tb))
