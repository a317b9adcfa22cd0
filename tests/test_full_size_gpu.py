"""Parity at BASELINE's full sizes (configs[2]: CelebA net Unet(256, 3, (1,2,2,2,2)), batch 1024) through size-independent
properties, since the fp32 CPU oracle cannot run 1024 samples of a 223 M-parameter net in seconds:
  * the score net is per-sample independent (GroupNorm only), so a few rows of the full-batch output are checked
    against the oracle run on just those samples;
  * one full predictor-corrector step at batch 1024: batch shards with the exact 2-scalar reduction reproduce the
    unsharded step, and CUDA-graph replay equals the eager loop."""
import pytest
import torch

from oracle import unet_oracle as uo
from oracle.det_weights import fill_state_dict
from tests.util import rel_l2, rel_max

pytestmark = pytest.mark.gpu
KW = dict(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2))


@pytest.fixture(scope="module")
def celeba_net():
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    m = Unet(**KW)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def test_full_batch_rows_match_oracle(celeba_net):
    m, sd = celeba_net
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(1024, 3, 16, 16, generator=g)
    t = torch.rand(1024, generator=g) * 0.999 + 1e-3
    with torch.no_grad():
        y = m(x.cuda(), t.cuda())
        pick = torch.tensor([0, 517, 1023])
        ref = uo.unet_forward(sd, x[pick], t[pick], dim=KW["dim"], dim_mults=KW["dim_mults"])
    assert y.shape == (1024, 3, 16, 16) and torch.isfinite(y).all()
    err = rel_l2(y[pick.cuda()], ref)
    print(f"CelebA net, batch 1024, rows {pick.tolist()}: rel-L2 vs oracle = {err:.3e}")
    assert err < 1.5e-2
    # the same rows computed alone take other kernels (single-CTA GEMM, other tile shapes): fp32 summation order moves
    # a few bf16 roundings, which a random-weight net amplifies to the size of its bf16 error (the oracle is the gate)
    with torch.no_grad():
        y_small = m(x[pick].cuda(), t[pick].cuda())
    e_small = rel_l2(y_small, y[pick.cuda()])
    print(f"rows alone vs inside the batch: {e_small:.3e}; alone vs oracle: {rel_l2(y_small, ref):.3e}")
    assert e_small < 1.2e-2 and rel_l2(y_small, ref) < 1.5e-2


def test_full_batch_pc_step_sharding_and_graph(celeba_net):
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    m, _ = celeba_net
    sde = sh.VPSDE(0.1, 20.0, 1000)
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1024, 3, 16, 16, generator=g).cuda()
    x0 = torch.randn(1024, 3, 16, 16, generator=g).cuda()
    sh.manual_seed(5)
    full = sh.cond_sampler(z, "0", "012", m, sde, x_init=x0, num_steps=2)
    sh.manual_seed(5)
    graph = sh.cond_sampler(z, "0", "012", m, sde, x_init=x0, num_steps=2, use_graph=True)
    assert rel_max(graph, full) < 1e-5
    assert torch.equal(full[:, 0], z[:, 0])  # observed modality returns its clean latent
    # two shards, exact mode, emulated on one GPU: record every shard's norm sums along the full trajectory is only
    # exact for the first corrector call, so compare a single step
    sh.manual_seed(5)
    one = sh.cond_sampler(z, "0", "012", m, sde, x_init=x0, num_steps=1)
    accs = {}

    def run(lo, hi, phase):
        sh.manual_seed(5, sample_offset=lo)
        calls = {"n": 0}

        def reduce_fn(acc):
            k = calls["n"]
            calls["n"] += 1
            if phase == 0:
                accs.setdefault(k, []).append(acc.clone())
            else:
                acc.copy_(sum(accs[k]))
        return sh.cond_sampler(z[lo:hi], "0", "012", m, sde, x_init=x0[lo:hi], num_steps=1, global_batch=1024,
                               reduce_fn=reduce_fn)
    run(0, 512, 0); run(512, 1024, 0)
    both = torch.cat([run(0, 512, 1), run(512, 1024, 1)])
    assert rel_l2(both, one) < 5e-3  # the net sees other tile shapes at batch 512 (bf16 rounding flips, see above)
