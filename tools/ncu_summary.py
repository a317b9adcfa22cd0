"""Extract the judged metrics of every profiled launch in an .ncu-rep into a small CSV (profiles/ evidence).
Usage: python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
kn = hdr.index("Kernel Name")
cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
with open(sys.argv[2], "w", newline="") as fh:
    wr = csv.writer(fh)
    wr.writerow(["kernel"] + [f"{w} [{units[i]}]" for w, i in cols])
    for r in data:
        wr.writerow([re.sub(r"\(.*", "", r[kn])[:80]] + [r[i] for _, i in cols])
print(f"{len(data)} launches -> {sys.argv[2]}")
