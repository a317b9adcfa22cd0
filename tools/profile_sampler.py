"""The fused sampler-step kernels (predictor, corrector norms, corrector update) and the DSM perturb / loss kernels
on the large-batch sweep point, bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`.
Usage: python tools/profile_sampler.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
sde = sh.VPSDE(1.0, 5.0, 100)
x = torch.randn(batch, 5, 8, 8, device=dev)
s = torch.randn_like(x)
t = torch.full((batch,), 0.5, device=dev)
rng = sh._RngState()
acc = torch.zeros(3, dtype=torch.float64, device=dev)
out = torch.empty_like(x)


def pc_kernels():
    # as the samplers run them: the Philox noise-norm kernel (no memory traffic) on a side stream, joined before the
    # score-norm kernel; set SBM_PROFILE_FUSED_NORMS=1 for the round-1 single norms kernel that regenerates the noise
    r_pred, r_corr = rng.next(), rng.next()
    if os.environ.get("SBM_PROFILE_FUSED_NORMS") == "1":
        x1, _ = sh._predictor_kernel(sde, x, s, t, rng=r_pred, want_mean=False, out=out)
        sh._corrector_kernels(sde, x1, s, t, 0.16, rng=r_corr, want_mean=False, acc=acc, out=out)
        return
    side = sh._fork_noise_norm(x, r_corr, acc)
    x1, _ = sh._predictor_kernel(sde, x, s, t, rng=r_pred, want_mean=False, out=out)
    sh._corrector_kernels(sde, x1, s, t, 0.16, rng=r_corr, want_mean=False, acc=acc, out=out, noise_norm_done=True,
                          join=side)


def dsm():
    return sh.loss_fn(x, lambda a, b: s, sde, likelihood_weighting=False)


with torch.no_grad():
    for _ in range(3):
        pc_kernels()
        dsm()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pc_kernels()
    dsm()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
