"""Pipeline trace of cluster 0 of the CTA-pair GEMM on a K-short layer (build the library with
SBM_NVCC_EXTRA=-DSBM_PAIR_TRACE first; the production build has no trace code).
python tools/trace_pair.py [cin] [cout] [k] [H] [batch]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import _lib as L, ops  # noqa: E402

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H = int(sys.argv[4]) if len(sys.argv) > 4 else 16
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
dev = torch.device("cuda")
x = torch.randn(B, H, H, ops.pad8(cin), device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
wpk = ops.pack_conv2d_weight(w)
bias = torch.randn(cout, device=dev)
outf = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.float32, device=dev)
for _ in range(3):
    ops.conv_igemm(x, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias, out=outf)
buf = torch.zeros(60001, dtype=torch.int64, device=dev)
L.lib().sbm_debug_pair_trace(C.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.conv_igemm(x, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias, out=outf)
e1.record()
torch.cuda.synchronize()
L.lib().sbm_debug_pair_trace(C.c_void_p(0))
print(f"launch {e0.elapsed_time(e1) * 1e3:.1f} us")
h = buf.cpu().tolist()
n = h[0]
recs = []
for v in h[1:1 + min(n, 60000)]:
    v &= (1 << 64) - 1
    recs.append((v >> 20, (v >> 19) & 1, (v >> 16) & 7, (v >> 12) & 15, v & 0xFFF))
t0 = min(r[0] for r in recs)
names = {(0, 0): "prod tile", (0, 1): "prod stage", (1, 0): "mma tile", (1, 1): "mma tempty", (1, 2): "mma full",
         (1, 3): "mma commit", (2, 0): "epi tile", (2, 1): "epi tfull", (2, 2): "epi release"}
recs.sort()
for t, cta, role, ev, rnd in recs:
    if rnd <= 6 or rnd >= 12 or True:
        print(f"{t - t0:8d} ns  cta{cta} round {rnd:3d}  {names.get((role, ev), (role, ev))}")
