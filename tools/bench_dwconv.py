"""Time the depthwise 7x7 kernel (bf16 output + statistics) at the CelebA score-net shapes, batch 1024, and check it
against float64 torch.  SBM_DWCONV_MMA=0 selects the FFMA2 kernel.  python tools/bench_dwconv.py [batch]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (H, C) in [(16, 170), (16, 256), (16, 512), (8, 256), (8, 512), (8, 1024), (4, 512), (4, 1024)]:
    g = torch.Generator().manual_seed(H * 1000 + C)
    ld = ops.pad8(C)
    x = torch.randn(B, H, H, ld, generator=g).to(dev)
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev)
    bias = torch.randn(C, generator=g).to(dev)
    cond = torch.randn(B, ld, generator=g).to(dev)
    ts = []
    for it in range(6):
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.dwconv7(x, C, w, bias, cond, ld, stats, out_dtype=torch.bfloat16)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    tt = sorted(ts[1:])[len(ts[1:]) // 2]
    nb = min(B, 64)
    xr = x[:nb, :, :, :C].permute(0, 3, 1, 2).double()
    ref = F.conv2d(xr, w.double(), bias.double(), padding=3, groups=C) + cond[:nb, :C].double()[:, :, None, None]
    got = out[:nb, :, :, :C].permute(0, 3, 1, 2).double()
    err = ((got - ref).norm() / ref.norm()).item()
    sref = torch.stack([got.sum(dim=(1, 2, 3)), (got * got).sum(dim=(1, 2, 3))], -1)
    serr = ((stats[:nb] - sref).abs().max() / sref.abs().max()).item()
    gb = B * H * H * C * 6 / 1e9
    print(f"H={H:2d} C={C:4d} B={B}: {tt:8.1f} us  {gb / tt * 1e3:6.2f} TB/s of 6 B/elt  rel-L2 {err:.2e}  stats {serr:.1e}  "
          f"(SBM_DWCONV_MMA={os.environ.get('SBM_DWCONV_MMA', '1')})", flush=True)
