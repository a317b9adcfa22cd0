"""One score-net forward (and optionally one full PC step) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python tools/profile_forward.py [celeba|poly] [batch] [pc]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402
from score_based_multimodal_autoencoder_b200.unet_model import Unet  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "celeba"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pc = len(sys.argv) > 3 and sys.argv[3] == "pc"
if which == "celeba":
    kw, shape, sde = dict(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)), (batch, 3, 16, 16), sh.VPSDE(0.1, 20.0, 1000)
else:
    kw, shape, sde = dict(dim=64, channels=5, dim_mults=(1, 2, 2, 2)), (batch, 5, 8, 8), sh.VPSDE(1.0, 5.0, 100)
torch.manual_seed(0)
m = Unet(**kw).cuda().eval()
x = torch.randn(*shape, device="cuda")
t = torch.full((batch,), 0.5, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x, t)
    if pc:
        sh.pc_sampler(x, m, sde, z_obs=x, obs_mask=1, num_steps=2)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if pc:
        sh.pc_sampler(x, m, sde, z_obs=x, obs_mask=1, num_steps=1)
    else:
        m(x, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
