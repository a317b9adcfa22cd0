"""Legacy evaluators: annealed-Langevin sampler (eval_lat_celeba_hq_all.py:258-275) and the fixed-step
Langevin refinement (fid_upd10.py:279-290) on the fused `sbm_langevin_axpy_step` kernel, against the oracle's restatement
of the two script loops with an exact fp32 score (<= 1e-5 per step, 1e-4 over the loop)."""
import numpy as np
import pytest
import torch

from oracle import sde_oracle as so
from tests.util import rel_max

pytestmark = pytest.mark.gpu


def _toy(x, idx):
    s = 1.0 + 0.01 * idx.float()
    return -x / s[:, None, None, None] + 0.1 * torch.sin(2.0 * x)


@pytest.mark.parametrize("given,n_comp,shape", [("0", 2, (7, 3, 16, 16)), ("12", 1, (5, 3, 16, 16)), ("", 1, (33, 5, 8, 8))])
def test_annealed_langevin_sampler_vs_oracle(given, n_comp, shape):
    from score_based_multimodal_autoencoder_b200 import eval_samplers as es
    B, M, D, _ = shape
    mods = "01234"[:M]
    g = torch.Generator().manual_seed(3)
    z = torch.randn(*shape, generator=g)
    sig = np.linspace(5, 0.1, 40)
    noise = torch.randn(len(sig), n_comp, *shape, generator=g)
    er = {m: 0.01 + 0.004 * i for i, m in enumerate(mods)}
    c = {m: 1.0 - 0.1 * i for i, m in enumerate(mods)}
    mask = [m in given for m in mods]
    ref = so.annealed_langevin(_toy, z, mask, [er[m] for m in mods], [c[m] for m in mods], sig, n_comp, noise)
    out = es.annealed_langevin_sampler(z.cuda(), given, mods, _toy, er, c, n_comp=n_comp, sigmas=sig, noise=noise.cuda())
    assert rel_max(out, ref) < 1e-4
    for i, on in enumerate(mask):
        if on:
            assert torch.equal(out[:, i].cpu(), z[:, i])
    # dict input of per-modality [B, size_z] latents (the reference's `z[mod]`), in-kernel Philox noise: runs, finite,
    # observed modalities untouched, reproducible under the same seed
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    zd = {m: z[:, i].reshape(B, -1).cuda() for i, m in enumerate(mods)}
    sh.manual_seed(9)
    a = es.annealed_langevin_sampler(zd, given, mods, _toy, er, c, n_comp=n_comp, sigmas=sig, num_levels=5)
    sh.manual_seed(9)
    b = es.annealed_langevin_sampler(zd, given, mods, _toy, er, c, n_comp=n_comp, sigmas=sig, num_levels=5)
    assert torch.equal(a, b) and torch.isfinite(a).all() and a.shape == shape


@pytest.mark.parametrize("schedule", [False, True])
def test_langevin_refine_vs_oracle(schedule):
    from score_based_multimodal_autoencoder_b200 import eval_samplers as es
    B, M, D = 9, 5, 8
    mods = "01234"
    g = torch.Generator().manual_seed(4)
    z = torch.randn(B, M, D, D, generator=g)
    w = torch.randn(M * D * D, M * D * D, generator=g) * 0.02
    n_comp = 6
    noise = torch.randn(n_comp + 1, B, M, D, D, generator=g)
    mask = [m not in "13" for m in mods]
    ref = so.langevin_refine(lambda f: f @ w, z, mask, n_comp, 0.05, 0.01, schedule, noise)
    out = es.langevin_refine(z.cuda(), "13", mods, lambda f: f @ w.cuda(), n_comp, 0.05, 0.01, schedule,
                             noise=noise.cuda())
    assert rel_max(out, ref) < 1e-4


def test_annealed_langevin_with_the_score_net_and_checkpoint_round_trip(tmp_path):
    """The evaluator drives this package's Unet with the integer level index as its time input, and the reference's
    checkpoint container round-trips through save_checkpoint / load_checkpoint."""
    from score_based_multimodal_autoencoder_b200 import eval_samplers as es
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    torch.manual_seed(0)
    m = Unet(dim=32, channels=3, dim_mults=(1, 2)).cuda().eval()
    p = tmp_path / "celeb_hq_cont_256_test"
    es.save_checkpoint(str(p), m, epoch=7, train_loss=0.5, val_loss=0.6, size_z=256)
    torch.manual_seed(1)
    m2 = Unet(dim=32, channels=3, dim_mults=(1, 2)).cuda().eval()
    meta = es.load_checkpoint(str(p), m2, map_location="cuda")
    assert meta == {"epoch": 7, "train_loss": 0.5, "val_loss": 0.6, "size_z": 256}
    z = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(2)).cuda()
    t = torch.rand(4).cuda()
    with torch.no_grad():
        assert torch.equal(m(z, t), m2(z, t))
    er = {k: 0.01 for k in "012"}
    c = {k: 1.0 for k in "012"}
    out = es.annealed_langevin_sampler(z, "0", "012", m2, er, c, n_comp=1, num_levels=6)
    assert out.shape == z.shape and torch.isfinite(out).all() and torch.equal(out[:, 0], z[:, 0])
