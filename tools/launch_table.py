"""Per-kernel table of an ncu launch list with several metrics per launch (gpu__time_duration.sum, dram__bytes_read/
write.sum, tensor-pipe activity): launches, total time, share, DRAM bytes, achieved DRAM rate, time-weighted tensor
activity; then the longest launches.  Usage: python tools/launch_table.py <launches.csv> [top]"""
import collections
import csv
import re
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)   # `| head` closes the pipe early

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = collections.OrderedDict()
for r in csv.DictReader([l for l in open(path) if l.startswith('"')]):
    d = rows.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])[:70], "grid": r.get("Grid Size", "")})
    v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    elif m.startswith("dram__bytes"):
        d["bytes"] = d.get("bytes", 0.0) + v * UNIT[u]
    elif "tensor" in m:
        d["tc"] = v
rows = list(rows.values())
tot = sum(r["us"] for r in rows)
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += r["us"]
    a[2] += r.get("bytes", 0.0)
    a[3] += r.get("tc", 0.0) * r["us"]
print(f"{path}: {tot / 1000:.3f} ms over {len(rows)} launches, DRAM {sum(r.get('bytes', 0) for r in rows) / 1e9:.2f} GB")
for k, (n, t, b, tc) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} n={n:3d} {t / 1000:8.3f} ms {100 * t / tot:5.1f}% {b / 1e9:7.2f} GB {b / t / 1e6:5.2f} TB/s tensor {tc / t:5.1f}%")
for r in sorted(rows, key=lambda r: -r["us"])[:top]:
    print(f"  {r['name']:62s} {r['us']:8.1f} us {r.get('bytes', 0) / 1e6:8.1f} MB tensor {r.get('tc', 0):5.1f}% grid {r['grid']}")
