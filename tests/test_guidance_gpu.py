"""Classifier / EBM guidance inside the samplers (sde_helper2.py:65-94, 283-312) on the CUDA path against
tests/golden/guidance.pt: outputs of the UNMODIFIED reference's em_predictor / corrector run with three pair energy nets as
`cl_g` (oracle/gen_golden_guidance.py).  The energy MLP and its input gradient run as bf16 tensor-core GEMMs, so the
guided score is bounded by rel-L2 <= 2e-2 of the fp32 reference (the bf16 bound of the score net itself) and the step
outputs, which add the guided score scaled by the step size to x, by 5e-3."""
import pytest
import torch

from oracle import guidance_oracle as go
from oracle.gen_golden_guidance import pair_weights
from tests.util import golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _nets(g, cls_name="ClwithTime2"):
    from score_based_multimodal_autoencoder_b200 import guidance as G
    nets, sds = {}, {}
    for pair in ("01", "02", "12"):
        sd = {k.split(".", 1)[1]: v for k, v in pair_weights(pair).items()}
        n = getattr(G, cls_name)(n_mod=2, size_z=g["size_z"], n_class=1, hidden=g["hidden"], time_dim=g["time_dim"])
        n.load_state_dict(sd, strict=(cls_name == "ClwithTime2"))
        nets[pair], sds[pair] = n.cuda().eval(), sd
    return nets, sds


def test_energy_gradient_kernels_vs_autograd():
    """d mean(E) / d x from the GEMM + act_bwd kernels vs torch.autograd through the fp32 functional energy."""
    g = golden("guidance.pt")
    nets, sds = _nets(g)
    gen = torch.Generator().manual_seed(8)
    for B in (6, 200):
        x = torch.randn(B, 2 * g["size_z"], generator=gen)
        t = torch.rand(B, generator=gen)
        ref = go.energy_grad(lambda f, tt: go.energy(sds["01"], f, tt, time_dim=g["time_dim"]), x, t)
        got = nets["01"].energy_grad(x.cuda(), t.cuda())
        err = rel_l2(got, ref)
        print(f"energy gradient, batch {B}: rel-L2 vs fp32 autograd = {err:.3e}")
        assert err < 2e-2
        # the module's own torch forward is the same function as the oracle's functional form
        out = nets["01"](x.cuda(), t.cuda())
        assert rel_max(out, go.energy(sds["01"], x, t, time_dim=g["time_dim"])) < 1e-4


def test_guided_predictor_and_corrector_match_reference_golden():
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    g = golden("guidance.pt")
    nets, _ = _nets(g)
    sde = sh.VPSDE(*g["sde"])
    x, t = g["x"].cuda(), g["t"].cuda()
    for c in g["cases"]:
        fn = lambda xx, tt: g["score"].cuda().clone()
        # the guided score itself (RSDE.sde's edit, sde_helper2.py:283-312)
        gs = sh._guided(g["score"].cuda().clone(), x, t, nets, g["cl_s"], c["given"], "012")
        e_s = rel_l2(gs, c["guided_score"])
        xp, xm = sh.em_predictor(x, t, fn, sde, cl_g=nets, cl_s=g["cl_s"], given=c["given"], all_mods="012",
                                 noise=g["z_pred"].cuda())
        xc, xcm = sh.corrector(x, t, fn, sde, 1, 0.16, cl_g=nets, cl_s=g["cl_s"], given=c["given"], all_mods="012",
                               noise=g["z_corr"].cuda())
        errs = [rel_l2(xp, c["pred_x"]), rel_l2(xm, c["pred_mean"]), rel_l2(xc, c["corr_x"]), rel_l2(xcm, c["corr_mean"])]
        print(f"given={c['given']!r}: guided score rel-L2 {e_s:.3e}; predictor / corrector outputs {max(errs):.3e}")
        assert e_s < 2e-2 and max(errs) < 5e-3
        # reverse SDE object: drift carries the guided score (sde_helper2.py:283-314)
        drift, diff = sde.reverse(fn).sde(x, t, nets, g["cl_s"], None, given=c["given"], all_mods="012")
        d0, _ = sde.sde(x, t)
        want = d0 - diff[:, None, None, None] ** 2 * c["guided_score"].cuda()
        assert rel_l2(drift, want) < 2e-2


def test_arbitrary_callable_takes_the_autograd_route_and_guided_sampler_runs():
    """`cl_g` values that are not this package's energy classes are differentiated with torch.autograd like the reference
    does; the N-step conditional sampler accepts the guidance arguments (train_lat_celebhq_unet_cont2.py:305-307)."""
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    g = golden("guidance.pt")
    nets, sds = _nets(g)
    x, t = g["x"].cuda(), g["t"].cuda()
    plain = {k: (lambda n: (lambda f, tt: n(f, tt)))(n) for k, n in nets.items()}      # hides the class: autograd route
    a = sh._guided(g["score"].cuda().clone(), x, t, plain, g["cl_s"], "0", "012")
    assert rel_l2(a, g["cases"][0]["guided_score"]) < 1e-4                                 # fp32 torch path
    b = sh._guided(g["score"].cuda().clone(), x, t, nets, g["cl_s"], "0", "012")
    assert rel_l2(b, a) < 2e-2
    toy = lambda xx, tt: -xx * (0.5 + tt[:, None, None, None])
    sde = sh.VPSDE(0.1, 20.0, 20)
    sh.manual_seed(3)
    plain_run = sh.cond_sampler(x, "0", "012", toy, sde, x_init=g["z_pred"].cuda(), num_steps=4)
    sh.manual_seed(3)
    guided = sh.cond_sampler(x, "0", "012", toy, sde, x_init=g["z_pred"].cuda(), num_steps=4, cl_g=nets, cl_s=g["cl_s"])
    assert torch.isfinite(guided).all() and torch.equal(guided[:, 0], x[:, 0])
    assert rel_l2(guided[:, 1:], plain_run[:, 1:]) > 1e-3                                  # guidance moved the sample
    # index-conditioned net (ClwithTime3, train_poly_unet_cont.py:72-88): predicted modality only
    from score_based_multimodal_autoencoder_b200.guidance import ClwithTime3
    torch.manual_seed(0)
    n3 = ClwithTime3(2, g["size_z"], 1, hidden=g["hidden"], time_dim=g["time_dim"]).cuda().eval()
    torch.manual_seed(1)
    s3 = sh._guided(g["score"].cuda().clone(), x, t, n3, g["cl_s"], "0", "012")
    changed = [(s3[:, i] - g["score"].cuda()[:, i]).abs().max().item() > 0 for i in range(3)]
    assert changed[0] is False and sum(changed[1:]) == 1
    m2 = changed.index(True)    # (the pair choice follows torch's host RNG, like the reference)
    new_x = torch.stack((x[:, 0], x[:, m2]), 1)
    ref = go.energy_grad(lambda f, tt: n3(f, tt, 0, m2), new_x, t)
    assert rel_l2(s3[:, m2], g["score"].cuda()[:, m2] - g["cl_s"] * ref[:, 1]) < 2e-2
