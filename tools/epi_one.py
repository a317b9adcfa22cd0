"""One K-short conv launch for ncu source-level profiling: python tools/epi_one.py [label]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import _lib as L, ops  # noqa: E402

B, H, cin, cout, k = 1024, 16, 128, 256, 1
dev = torch.device("cuda")
x = torch.randn(B, H, H, ops.pad8(cin), device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
wpk = ops.pack_conv2d_weight(w)
bias = torch.randn(cout, device=dev)
outf = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.float32, device=dev)
for _ in range(4):
    ops.conv_igemm(x, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias, out=outf)
torch.cuda.synchronize()
