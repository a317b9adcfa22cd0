"""Time individual sbm_conv_igemm launches (CUDA events on the launch stream, warm, L2 flushed between launches) for
the layer shapes of the CelebA score net at batch 1024.  Usage: python tools/bench_conv_shapes.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import _lib as L, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda")
SHAPES = [  # H, cin, cout, k
    (16, 512, 256, 3), (16, 256, 512, 3), (8, 1024, 512, 3), (8, 512, 1024, 3), (4, 1024, 512, 3), (4, 512, 1024, 3),
    (2, 1024, 512, 3), (2, 512, 1024, 3), (2, 1024, 1024, 3), (1, 1024, 512, 3), (1, 512, 1024, 3), (1, 1024, 1024, 3),
    (1, 1024, 384, 1), (1, 128, 1024, 1), (16, 147, 170, 1), (16, 170, 256, 1),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
MODES = [("staged", 0)] + ([("direct", 1)] if os.environ.get("SBM_AB_EPILOGUE") else [])
if os.environ.get("SBM_LOWRES"):
    SHAPES = [(2, 1024, 512, 3), (2, 512, 1024, 3), (2, 1024, 1024, 3), (2, 512, 512, 3), (4, 1024, 512, 3), (4, 512, 1024, 3),
              (1, 1024, 512, 3), (1, 512, 1024, 3), (2, 512, 384, 1), (4, 512, 384, 1)]
if os.environ.get("SBM_SMALLK"):
    SHAPES = [(16, 147, 170, 1), (16, 170, 256, 1), (16, 256, 384, 1), (16, 128, 256, 1), (8, 512, 384, 1), (8, 128, 512, 1),
              (16, 512, 256, 1), (16, 256, 512, 3)]
for (H, cin, cout, k) in SHAPES:
    x = torch.randn(B, H, H, ops.pad8(cin), device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    wpk = ops.pack_conv2d_weight(w)
    bias = torch.randn(cout, device=dev)
    out = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.bfloat16, device=dev)
    st = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    taps = sum(1 for i in range(k) for j in range(k) if abs(i - k // 2) < H and abs(j - k // 2) < H)
    flops = 2.0 * B * H * H * cin * cout * taps
    outf = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.float32, device=dev)
    for mode, flag in MODES:
        L.lib().sbm_conv_force_direct_epilogue(flag)
        labels = [("bf16+gelu+stats", dict(act=L.ACT_GELU, out=out, stats=st)), ("fp32+bf16copy", dict(out=outf, out2=out))]
        if os.environ.get("SBM_STATS_AB"):
            labels = [("bf16+gelu+stats", dict(act=L.ACT_GELU, out=out, stats=st)), ("bf16+gelu", dict(act=L.ACT_GELU, out=out)),
                      ("bf16+stats", dict(out=out, stats=st)), ("bf16", dict(out=out))]
        for label, kw in labels:
            ts = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv_igemm(x, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias, **kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            t = sorted(ts[1:])[len(ts[1:]) // 2]
            v = L.lib().sbm_conv_last_variant()
            print(f"H={H:2d} {cin:4d}->{cout:4d} k={k} M={B * H * H:6d} taps={taps} {mode:6s} {label:16s}: {t:8.1f} us  "
                  f"{flops / t / 1e6:7.1f} TFLOP/s  BN={v & 0xFFFF}{' pair' if v & (1 << 16) else ''}", flush=True)
    L.lib().sbm_conv_force_direct_epilogue(0)
