"""Time eager + graphed DSM training steps of the CelebA net: python tools/time_train.py [batch] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402
from score_based_multimodal_autoencoder_b200.optim import GraphedTrainStep  # noqa: E402
from score_based_multimodal_autoencoder_b200.unet_model import Unet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.manual_seed(0)
m = Unet(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)).cuda().train()
sde = sh.VPSDE(0.1, 20.0, 1000)
x = torch.randn(B, 3, 16, 16, device="cuda")
step = GraphedTrainStep(m, sde, x, lr=5e-5, warmup=3)
for _ in range(2):
    step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    loss = step(x)
e1.record()
torch.cuda.synchronize()
print(f"train step B={B}: {e0.elapsed_time(e1) / iters:.3f} ms  loss {loss.item():.4f}  "
      f"(SBM_WGRAD_ATOMIC={os.environ.get('SBM_WGRAD_ATOMIC', '0')})")
step.close()
