"""Build libsbmae_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

`python -m score_based_multimodal_autoencoder_b200.build` or `__graft_entry__.build()`.
nvcc cross-compiles for sm_100a without a GPU; the .so travels to the GPU box with the tree.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libsbmae_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
] + os.environ.get("SBM_NVCC_EXTRA", "").split()   # e.g. -DSBM_PAIR_TRACE for tools/trace_pair.py


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libsbmae_b200 cannot be built")
    return cand


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths: list[str]) -> str:
    h = hashlib.sha256()
    hdrs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "sbmae_b200.h"))
    for p in paths + hdrs:
        with open(p, "rb") as fh:
            h.update(p.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp.txt")
    dig = _digest(srcs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> tuple[str, str]:
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    with open(os.path.join(BUILD, "ptxas.log"), "w") as fh:
        for _, log in results:
            fh.write(log)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(f"built {LIB} from {len(srcs)} sources")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
