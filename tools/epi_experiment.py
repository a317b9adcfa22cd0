"""A/B of the statically compiled epilogue loops of the CTA-pair GEMM against the run-time tested loop (EM_DYN) on
K-short and K-long layers.  python tools/epi_experiment.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import _lib as L, ops  # noqa: E402

B = 1024
dev = torch.device("cuda")
SHAPES = [(16, 128, 256, 1), (16, 256, 512, 3), (16, 512, 256, 3), (8, 128, 512, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
DBG = [1, 0]   # SBM_EPI_STATIC: statically compiled epilogue loop vs the run-time tested one
for (H, cin, cout, k) in SHAPES:
    x = torch.randn(B, H, H, ops.pad8(cin), device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    wpk = ops.pack_conv2d_weight(w)
    bias = torch.randn(cout, device=dev)
    out = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.bfloat16, device=dev)
    outf = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.float32, device=dev)
    res = torch.randn(B, H, H, ops.pad8(cout), dtype=torch.float32, device=dev)
    st = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    labels = [("bf16", dict(out=out)), ("bf16+gelu+stats", dict(act=L.ACT_GELU, out=out, stats=st)),
              ("fp32", dict(out=outf)), ("fp32+bf16copy", dict(out=outf, out2=out)),
              ("fp32+res+copy", dict(out=outf, out2=out, residual=res))]
    for label, kw in labels:
        for dbg in DBG:
            L.lib().sbm_conv_epilogue_static(dbg)
            ts = []
            for it in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv_igemm(x, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias, **kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            t = sorted(ts[1:])[len(ts[1:]) // 2]
            print(f"H={H:2d} {cin:4d}->{cout:4d} k={k} {label:16s} static={dbg}: {t:8.1f} us", flush=True)
L.lib().sbm_conv_epilogue_static(1)
