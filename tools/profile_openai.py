"""One eval-mode forward of the z-conditioned UNetModel (batch 1024) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python tools/profile_openai.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200.unet_openai import UNetModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = UNetModel(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(), dropout=0.1,
              channel_mult=(1, 2, 4, 8), num_heads=1, use_z=True, z_dim=512).cuda().eval()
with torch.no_grad():
    for p in m.parameters():
        if p.abs().max() == 0:
            p.normal_(0, 0.02)
    x = torch.randn(B, 3, 16, 16, device="cuda")
    t = torch.rand(B, device="cuda") * 999
    z = torch.randn(B, 512, device="cuda")
    for _ in range(2):
        m(x, t, z)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x, t, z)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
