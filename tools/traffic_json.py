"""Aggregate an ncu launch list (gpu__time_duration.sum, dram__bytes_read/write.sum, tensor-pipe activity; --csv
log of `ncu --metrics ... python tools/profile_forward.py celeba 1024`) per kernel into the JSON `bench.py` reads
`roofline.traffic` from.  All epilogue modes of conv_igemm_pair_kernel<256,5,staged,MODE> are one kernel family (same
main loop, the 4th template argument only selects the compiled epilogue flag set).
Usage: python tools/traffic_json.py profiles/<launches>.csv profiles/<out>.json"""
import collections
import csv
import io
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(io.StringIO("".join(l for l in open(src) if l.startswith('"')))))
launch = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("(int)", "").replace("(bool)", "").replace("(unsigned int)", "")
    launch.setdefault(r["ID"], {"k": name})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))


def summarise(items):
    us = [d.get("gpu__time_duration.sum", 0.0) / 1e3 for d in items]
    by = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in items]
    tp = [d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for d in items]
    tot = sum(us)
    return {"launches": len(items), "total_us": tot, "avg_dram_bytes_per_launch": sum(by) / max(1, len(items)),
            "tensor_pipe_active_pct_time_weighted": sum(u * t for u, t in zip(us, tp)) / tot if tot else 0.0}


per = collections.OrderedDict()
for d in launch.values():
    per.setdefault(d["k"], []).append(d)
# (the split-K GEMM passes -- epilogue mode 131072 = EM_SPLITK -- are left out like in bench.py's dominant-kernel filter)
dom = [d for d in launch.values() if re.search(r"conv_igemm_pair_kernel<256, 5, (1|true)", d["k"])
       and not re.search(r", 131072u?>", d["k"])]
out = {"conv_igemm_pair_kernel<256,5,staged>": summarise(dom), "source": f"{src} (ncu, one forward, batch 1024)",
       "forward_ms_serialised": sum(d.get("gpu__time_duration.sum", 0.0) for d in launch.values()) / 1e6,
       "per_kernel": {k: summarise(v) for k, v in sorted(per.items(), key=lambda kv: -sum(x.get("gpu__time_duration.sum", 0) for x in kv[1]))}}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["conv_igemm_pair_kernel<256,5,staged>"]), out["forward_ms_serialised"])
