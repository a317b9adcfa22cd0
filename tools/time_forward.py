"""Time the CelebA score-net forward (CUDA events, warm, graph-free): python tools/time_forward.py [batch] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200.unet_model import Unet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
m = Unet(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)).cuda().eval()
x = torch.randn(B, 3, 16, 16, device="cuda")
t = torch.rand(B, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x, t)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m(x, t)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        y = m(x, t)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
knobs = " ".join(k + "=" + v for k, v in os.environ.items() if k.startswith("SBM_")) or "defaults"
print(f"forward B={B}: {e0.elapsed_time(e1) / iters:.3f} ms  ({knobs})")
