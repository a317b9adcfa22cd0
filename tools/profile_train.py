"""One eager DSM training step (loss_fn + backward + FusedAdam) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python tools/profile_train.py [celeba|poly] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402
from score_based_multimodal_autoencoder_b200.optim import FusedAdam  # noqa: E402
from score_based_multimodal_autoencoder_b200.unet_model import Unet  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "celeba"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if which == "celeba":
    kw, shape, sde, lr = dict(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)), (batch, 3, 16, 16), sh.VPSDE(0.1, 20.0, 1000), 5e-5
else:
    kw, shape, sde, lr = dict(dim=64, channels=5, dim_mults=(1, 2, 2, 2)), (batch, 5, 8, 8), sh.VPSDE(1.0, 5.0, 100), 5e-4
torch.manual_seed(0)
m = Unet(**kw).cuda().train()
opt = FusedAdam(m.parameters(), lr=lr)
x = torch.randn(*shape, device="cuda")


def step():
    loss = sh.loss_fn(x, m, sde, reduce_mean=True, likelihood_weighting=False, eps=1e-5, rng="philox")
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
