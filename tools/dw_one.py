"""A few depthwise-conv launches for ncu: python tools/dw_one.py [H] [C]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import ops  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
B = 1024
dev = torch.device("cuda")
ld = ops.pad8(C)
x = torch.randn(B, H, H, ld, device=dev)
w = torch.randn(C, 1, 7, 7, device=dev) / 7
bias = torch.randn(C, device=dev)
cond = torch.randn(B, ld, device=dev)
for _ in range(4):
    stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    ops.dwconv7(x, C, w, bias, cond, ld, stats, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
