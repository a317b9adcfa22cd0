"""SASS evidence for the tcgen05 / TMA kernels of libsbmae_b200.so (run in the build container, no GPU needed):
    python tools/sass_census.py > profiles/r2_sass_census.txt
Per kernel: counts of the Blackwell mnemonics (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; LDTM = tcgen05.ld; UTCBAR =
tcgen05.commit; UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add; SYNCS = mbarrier; UCGABAR = cluster
barrier; FFMA2 = packed fp32), then the instruction lines themselves for the dominant kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "score_based_multimodal_autoencoder_b200", "csrc", "libsbmae_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.[A-Z0-9.]+)?|LDTM(?:\.[A-Za-z0-9.]+)?|UTMALDG(?:\.[A-Z0-9.]+)?|"
                 r"UTMASTG(?:\.[A-Z0-9.]+)?|UTMAREDG(?:\.[A-Z0-9.]+)?|UTMAPF(?:\.[A-Z0-9.]+)?|SYNCS(?:\.[A-Z0-9.]+)?|"
                 r"UCGABAR_ARV|UCGABAR_WAIT|FFMA2|HMMA(?:\.[A-Z0-9.]+)?|LDGSTS(?:\.[A-Z0-9.]+)?)\b")
fn, counts, lines = None, collections.OrderedDict(), collections.defaultdict(list)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        ops = pat.findall(line)
        for op in ops:
            counts[fn][op] += 1
        if ops and any(o.startswith(("UTCHMMA", "UTMA", "LDTM", "UTCBAR")) for o in ops):
            lines[fn].append(re.sub(r"\s+/\*[0-9a-fx]+\*/\s*$", "", line.strip()))


def demangle(n):
    return re.sub(r"\(.*", "", subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip())


print(__doc__.strip().splitlines()[0])
print(f"library: {os.path.relpath(LIB, ROOT)}  ({len(counts)} kernels)\n")
print("== kernels that issue tcgen05 / TMA instructions")
tc = [f for f, c in counts.items() if any(k.startswith(("UTCHMMA", "UTMA", "LDTM")) for k in c)]
for f in tc:
    print(demangle(f))
    print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(counts[f].items())))
print("\n== kernels using packed fp32 FFMA2 or cp.async (LDGSTS)")
for f, c in counts.items():
    if f not in tc and (c.get("FFMA2") or any(k.startswith("LDGSTS") for k in c)):
        print(demangle(f) + ": " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
dom = [f for f in tc if "conv_igemm_pair_kernelILi256ELi5ELb1" in f]
if dom:
    print(f"\n== tcgen05 / TMA instruction lines of the dominant kernel {demangle(dom[0])}")
    for ln in lines[dom[0]]:
        print("    " + ln)
