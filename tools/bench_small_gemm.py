"""Latency anatomy of SUB-WAVE implicit-GEMM launches (the 4x4 / 2x2 / 1x1 levels at small per-GPU batch, the PolyMNIST
net at batch 64): a chain of 20 identical launches replayed from one CUDA graph, per shape, so the figure is the
steady-state cost of one launch inside a captured step.  Usage: python tools/bench_small_gemm.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import _lib as L, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
SHAPES = [  # H, cin, cout, k
    (4, 256, 512, 3), (4, 512, 256, 3), (4, 2304, 512, 1), (4, 64, 512, 3), (4, 256, 64, 3),
    (2, 512, 512, 3), (2, 512, 256, 3), (1, 512, 512, 3), (1, 512, 1024, 1), (1, 1024, 1024, 1),
    (8, 256, 512, 3), (8, 512, 256, 3), (16, 256, 256, 3),
]
for (H, cin, cout, k) in SHAPES:
    x = torch.randn(B, H, H, ops.pad8(cin), device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    wpk = ops.pack_conv2d_weight(w)
    bias = torch.randn(cout, device=dev)
    out = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.bfloat16, device=dev)
    taps = sum(1 for i in range(k) for j in range(k) if abs(i - k // 2) < H and abs(j - k // 2) < H)
    kblocks = taps * ((cin + 63) // 64)
    flops = 2.0 * B * H * H * cin * cout * taps

    def run():
        ops.conv_igemm(x, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias, out=out)

    for _ in range(3):
        run()
    v = L.lib().sbm_conv_last_variant()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / 100
    bn = v & 0xFFFF
    m_tiles = (B * H * H + 127) // 128
    ctas = m_tiles * ((cout + bn - 1) // bn)
    print(f"B={B} H={H:2d} {cin:4d}->{cout:4d} k={k} M={B * H * H:6d} k-blocks={kblocks:3d} BN={bn}{' pair' if v & (1 << 16) else ''}"
          f"{' pm' if v & (1 << 18) else ''} CTAs~{ctas:4d}: {t:7.2f} us/launch  {t / kblocks * 1e3:6.0f} ns/k-block  "
          f"{flops / t / 1e6:7.1f} TFLOP/s", flush=True)
