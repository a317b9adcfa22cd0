"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
lines = [l for l in open(path) if not l.startswith("==")]
agg, allr, tot = collections.OrderedDict(), [], 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000.0 if unit == "ns" else (v * 1000.0 if unit == "ms" else v)
    short = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    allr.append((short, v, row.get("Grid Size", ""), row.get("Block Size", "")))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} n={n:4d} total={t / 1000:9.3f} ms {100 * t / tot:5.1f}%")
print(f"total {tot / 1000:.3f} ms over {len(allr)} launches")
for s, v, g, b in sorted(allr, key=lambda x: -x[1])[:top]:
    print(f"  {s:56s} {v:9.1f} us grid={g} block={b}")
