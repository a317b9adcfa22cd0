"""GPU parity of the fused sampler / DSM kernels (csrc/sampler.cu) and of the sampler entry points.

Tolerances: <= 1e-5 relative for the fp32 step kernels given the same score and noise (north_star asks <= 1e-3
per sampler step); the N-step loops with the bf16 score net are bounded by the stated drift bound 2e-2 rel-L2
(SURVEY.md Appendix E)."""
import pytest
import torch

from oracle import sde_oracle as so
from oracle import unet_oracle as uo
from oracle.det_weights import fill_state_dict
from tests.util import golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu
STEP_TOL = 1e-5
DRIFT_TOL = 2e-2


def _sh():
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    return sh


def _mk(kind, a, b, N):
    sh = _sh()
    return {"vp": sh.VPSDE, "subvp": sh.subVPSDE, "ve": sh.VESDE}[kind](a, b, N)


def test_step_kernels_match_reference_golden():
    sh = _sh()
    g = golden("sampler_steps.pt")
    sde = _mk(*g["sde"])
    x, t, score = g["x"].cuda(), g["t"].cuda(), g["score"].cuda()
    fn = lambda xx, tt: score
    xp, xm = sh.em_predictor(x, t, fn, sde, noise=g["z_pred"].cuda())
    assert rel_max(xp, g["pred_x"]) < STEP_TOL and rel_max(xm, g["pred_mean"]) < STEP_TOL
    xc, xcm = sh.corrector(x, t, fn, sde, 1, g["target_snr"], noise=g["z_corr"].cuda())
    assert rel_max(xc, g["corr_x"]) < STEP_TOL and rel_max(xcm, g["corr_mean"]) < STEP_TOL


@pytest.mark.parametrize("kind,a,b", [("vp", 0.1, 20.0), ("subvp", 0.1, 20.0), ("ve", 0.01, 50.0)])
def test_step_kernels_all_sdes_vs_oracle(kind, a, b):
    sh = _sh()
    N = 50
    sde = _mk(kind, a, b, N)
    spec = so.SdeSpec(kind, a, b, N)
    g = torch.Generator().manual_seed(3)
    B = 33
    x = torch.randn(B, 3, 16, 16, generator=g)
    t = torch.rand(B, generator=g) * 0.999 + 1e-3
    score = torch.randn(B, 3, 16, 16, generator=g)
    z = torch.randn(B, 3, 16, 16, generator=g)
    fn = lambda xx, tt: score.cuda()
    xp, xm = sh.em_predictor(x.cuda(), t.cuda(), fn, sde, noise=z.cuda())
    rp, rm = so.em_predictor_step(spec, x, t, score, z)
    assert rel_max(xp, rp) < STEP_TOL and rel_max(xm, rm) < STEP_TOL
    xo, xom = sh.em_predictor(x.cuda(), t.cuda(), fn, sde, probability_flow=True)
    ro, rom = so.em_predictor_step(spec, x, t, score, z, probability_flow=True)
    assert rel_max(xo, ro) < STEP_TOL and rel_max(xom, rom) < STEP_TOL and torch.equal(xo, xom)
    xc, xcm = sh.corrector(x.cuda(), t.cuda(), fn, sde, 1, 0.16, noise=z.cuda())
    rc, rcm = so.corrector_step(spec, x, t, score, z, 0.16)
    assert rel_max(xc, rc) < STEP_TOL and rel_max(xcm, rcm) < STEP_TOL


def test_reverse_diffusion_predictor_matches_reference_golden():
    """sbm_rd_predictor_step (a mode of the fused predictor kernel) vs the update built on the unmodified reference's
    sde.reverse(score_fn, pf).discretize (tests/golden/rd_predictor.pt): VP (two tables), subVP, VE; SDE and ODE."""
    sh = _sh()
    for c in golden("rd_predictor.pt"):
        sde = _mk(c["kind"], c["a"], c["b"], c["N"])
        score = c["score"].cuda()
        fn = lambda xx, tt: score
        xn, xm = sh.rd_predictor(c["x"].cuda(), c["t"].cuda(), fn, sde, noise=c["z"].cuda())
        assert rel_max(xn, c["sde"]["x"]) < STEP_TOL and rel_max(xm, c["sde"]["x_mean"]) < STEP_TOL, c["kind"]
        xo, xom = sh.rd_predictor(c["x"].cuda(), c["t"].cuda(), fn, sde, probability_flow=True)
        assert rel_max(xo, c["ode"]["x"]) < STEP_TOL and rel_max(xom, c["ode"]["x_mean"]) < STEP_TOL, c["kind"]
        assert torch.equal(xo, xom)


@pytest.mark.parametrize("kind,a,b", [("vp", 0.1, 20.0), ("subvp", 0.1, 20.0), ("ve", 0.01, 50.0)])
def test_reverse_diffusion_predictor_all_sdes_vs_oracle(kind, a, b):
    sh = _sh()
    N = 50
    sde = _mk(kind, a, b, N)
    spec = so.SdeSpec(kind, a, b, N)
    g = torch.Generator().manual_seed(4)
    B = 33
    x = torch.randn(B, 3, 16, 16, generator=g)
    t = torch.rand(B, generator=g) * 0.999 + 1e-3
    score = torch.randn(B, 3, 16, 16, generator=g)
    z = torch.randn(B, 3, 16, 16, generator=g)
    fn = lambda xx, tt: score.cuda()
    xp, xm = sh.rd_predictor(x.cuda(), t.cuda(), fn, sde, noise=z.cuda())
    rp, rm = so.rd_predictor_step(spec, x, t, score, z)
    assert rel_max(xp, rp) < STEP_TOL and rel_max(xm, rm) < STEP_TOL


def _toy_score(x, t):
    # smooth, per-sample, batch-independent stand-in for the net: isolates the sampler arithmetic
    return -x * (0.5 + t[:, None, None, None]) + 0.1 * torch.sin(3.0 * x)


@pytest.mark.parametrize("given,pf,nobs", [("0", True, True), ("13", True, True), ("0", False, True),
                                           ("24", True, False), ("", True, True)])
def test_pc_sampler_loop_logic_vs_oracle(given, pf, nobs):
    """Whole conditional loop (imputation epilogue, both orders, finishing rule) with an exact fp32 score."""
    sh = _sh()
    N, B, M, D = 12, 9, 5, 8
    sde = _mk("vp", 1.0, 5.0, N)
    spec = so.SdeSpec("vp", 1.0, 5.0, N)
    g = torch.Generator().manual_seed(11)
    z0 = torch.randn(B, M, D, D, generator=g)
    npred = torch.randn(N, B, M, D, D, generator=g)
    ncorr = torch.randn(N, 1, B, M, D, D, generator=g)
    mods = "01234"
    mask = [m in given for m in mods]
    ref = so.pc_sampler(spec, _toy_score, z0, npred, ncorr, z_obs=z0, obs_mask=mask, noise_obs=nobs,
                        predictor_first=pf)
    out = sh.cond_sampler(z0.cuda(), given, mods, _toy_score, sde, noise_obs=nobs,
                          pc_order="predictor_first" if pf else "corrector_first", x_init=z0.cuda(),
                          noise_pred=npred.cuda(), noise_corr=ncorr.cuda())
    assert rel_max(out, ref) < 1e-4
    for i, on in enumerate(mask):
        if on:
            assert torch.equal(out[:, i].cpu(), z0[:, i])


@pytest.mark.parametrize("kind,a,b,pf", [("vp", 1.0, 5.0, True), ("vp", 0.1, 20.0, False), ("ve", 0.01, 2.0, True),
                                         ("subvp", 0.1, 20.0, True)])
def test_pc_sampler_reverse_diffusion_loop_vs_oracle(kind, a, b, pf):
    """predictor="reverse_diffusion" selected in the N-step conditional sampler (imputation epilogue included)."""
    sh = _sh()
    # the DDPM rule needs beta_max / N < 1 (alpha_i = 1 - beta_i under a square root): N = 60 for beta_max = 20
    N, B, M, D = (12 if b <= 5.0 else 60), 9, 5, 8
    sde = _mk(kind, a, b, N)
    spec = so.SdeSpec(kind, a, b, N)
    g = torch.Generator().manual_seed(12)
    z0 = torch.randn(B, M, D, D, generator=g)
    npred = torch.randn(N, B, M, D, D, generator=g)
    ncorr = torch.randn(N, 1, B, M, D, D, generator=g)
    mask = [m in "02" for m in "01234"]
    ref = so.pc_sampler(spec, _toy_score, z0, npred, ncorr, z_obs=z0, obs_mask=mask, predictor_first=pf,
                        predictor="reverse_diffusion")
    out = sh.cond_sampler(z0.cuda(), "02", "01234", _toy_score, sde, x_init=z0.cuda(), noise_pred=npred.cuda(),
                          pc_order="predictor_first" if pf else "corrector_first", noise_corr=ncorr.cuda(),
                          predictor="reverse_diffusion")
    assert torch.isfinite(ref).all()
    assert rel_max(out, ref) < 1e-4
    # and through the cached CUDA graph with the in-kernel Philox stream: graph == eager
    sh.manual_seed(5)
    e = sh.cond_sampler(z0.cuda(), "02", "01234", _toy_score, sde, x_init=z0.cuda(), predictor="reverse_diffusion")
    sh.manual_seed(5)
    gr = sh.cond_sampler(z0.cuda(), "02", "01234", _toy_score, sde, x_init=z0.cuda(), predictor="reverse_diffusion",
                         use_graph=True)
    assert rel_max(gr, e) < 1e-5


def test_cond_loops_with_bf16_net_vs_reference_golden():
    sh = _sh()
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    g = golden("sampler_steps.pt")
    net = golden("unet_poly.pt")
    m = Unet(**net["kwargs"])
    m.load_state_dict(fill_state_dict(net["shapes"]))
    m = m.cuda().eval()
    sde = _mk(*g["sde"])
    z0 = g["loop_z0"].cuda()
    for loop in g["loops"]:
        out = sh.cond_sampler(z0, loop["given"], g["mods"], m, sde, noise_obs=loop["noise_obs"],
                              pc_order="predictor_first" if loop["predictor_first"] else "corrector_first",
                              x_init=z0, noise_pred=g["loop_npred"].cuda(), noise_corr=g["loop_ncorr"].cuda(),
                              num_steps=g["loop_steps"])
        err = rel_l2(out, loop["out"])
        print(f"given={loop['given']!r} pf={loop['predictor_first']}: drift after {g['loop_steps']} steps = {err:.3e}")
        assert err < DRIFT_TOL
    # library uncond_sampler semantics (corrector -> predictor), N = 10 steps
    out = sh.pc_sampler(g["uncond_prior"].cuda(), m, sde, predictor_first=False, noise_pred=g["uncond_npred"].cuda(),
                        noise_corr=g["uncond_ncorr"].cuda())
    err = rel_l2(out, g["uncond_out"])
    print(f"uncond N=10 drift = {err:.3e}")
    assert err < DRIFT_TOL


def test_philox_statistics_and_shard_invariance():
    sh = _sh()
    sh.manual_seed(123)
    a = sh.randn((64, 5, 8, 8), "cuda")
    assert abs(a.mean().item()) < 0.02 and abs(a.std().item() - 1.0) < 0.02
    # kurtosis of a normal = 3
    assert abs((a ** 4).mean().item() - 3.0) < 0.15
    # the second half drawn as a shard (sample_offset = 32) equals the tail of the full draw
    sh.manual_seed(123, sample_offset=32)
    b = sh.randn((32, 5, 8, 8), "cuda")
    assert torch.equal(a[32:], b)
    sh.manual_seed(123)
    c = sh.randn((64, 5, 8, 8), "cuda")
    d = sh.randn((64, 5, 8, 8), "cuda")
    assert torch.equal(a, c) and not torch.equal(c, d)


@pytest.mark.parametrize("pf,shape", [(True, (16, 5, 8, 8)), (False, (9, 3, 16, 16)), (True, (7, 1, 8, 8))])
def test_in_kernel_noise_norms_equal_the_materialised_philox_draws(pf, shape):
    """The corrector's noise norm never reads a noise tensor on the Philox path: the side-stream sbm_noise_norm kernel
    re-derives it from the counter-based stream (sum of squares of a Box-Muller pair = -2 ln u, no normals formed).  It
    must equal the norm of the draws themselves: run the sampler once on the in-kernel stream and once with the SAME
    draws materialised by sbm_randn and injected (that path reduces the noise tensor it is given)."""
    import ctypes as C
    from score_based_multimodal_autoencoder_b200 import _lib as L
    sh = _sh()
    N = 6
    sde = _mk("vp", 1.0, 5.0, N)
    g = torch.Generator().manual_seed(21)
    z = torch.randn(*shape, generator=g).cuda()
    x0 = torch.randn(*shape, generator=g).cuda()
    mods = "01234"[:shape[1]]
    order = "predictor_first" if pf else "corrector_first"
    sh.manual_seed(1234)
    out_p = sh.cond_sampler(z, "0" if shape[1] > 1 else "", mods, _toy_score, sde, x_init=x0, pc_order=order)

    def draw(d):
        t = torch.empty(shape, device="cuda")
        L.check(L.lib().sbm_randn(L.ptr(t), C.c_int64(t.numel()), C.c_uint64(1234), C.c_uint64(d), C.c_uint64(0),
                                  C.c_float(1.0), L.stream_ptr()), "sbm_randn")
        return t
    # draw ids in call order: predictor_first -> (pred, corr) per step; corrector_first -> (corr, pred)
    npred = torch.stack([draw(2 * i + (0 if pf else 1)) for i in range(N)])
    ncorr = torch.stack([draw(2 * i + (1 if pf else 0)) for i in range(N)])[:, None]
    out_i = sh.cond_sampler(z, "0" if shape[1] > 1 else "", mods, _toy_score, sde, x_init=x0, pc_order=order,
                            noise_pred=npred, noise_corr=ncorr)
    assert rel_max(out_p, out_i) < 1e-5


def test_graph_replay_equals_eager_and_sharding_is_exact():
    sh = _sh()
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    m = Unet(dim=32, channels=5, dim_mults=(1, 2, 2, 2)).cuda().eval()
    sde = _mk("vp", 1.0, 5.0, 8)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(16, 5, 8, 8, generator=g).cuda()
    x0 = torch.randn(16, 5, 8, 8, generator=g).cuda()
    sh.manual_seed(77)
    eager = sh.cond_sampler(z, "0", "01234", m, sde, x_init=x0)
    sh.manual_seed(77)
    graph = sh.cond_sampler(z, "0", "01234", m, sde, x_init=x0, use_graph=True)
    assert rel_max(graph, eager) < 1e-5
    # two shards with the exact 2-scalar reduction reproduce the full batch (SURVEY.md 8e mode 1): emulate the
    # all-reduce by summing the per-shard norm accumulators by hand
    accs = {}

    def run_shard(lo, hi, phase):
        sh.manual_seed(77, sample_offset=lo)
        calls = {"n": 0}

        def reduce_fn(acc):
            k = calls["n"]
            calls["n"] += 1
            if phase == 0:
                accs.setdefault(k, []).append(acc.clone())
            else:
                acc.copy_(sum(accs[k]))
        return sh.cond_sampler(z[lo:hi], "0", "01234", m, sde, x_init=x0[lo:hi], global_batch=16, reduce_fn=reduce_fn)

    # phase 0 records each shard's norms along ITS OWN trajectory, which is only exact for the first corrector
    # call; so compare a single-step run
    sde1 = _mk("vp", 1.0, 5.0, 8)
    def one(lo, hi, phase):
        sh.manual_seed(77, sample_offset=lo)
        calls = {"n": 0}
        def reduce_fn(acc):
            k = calls["n"]; calls["n"] += 1
            if phase == 0:
                accs.setdefault(k, []).append(acc.clone())
            else:
                acc.copy_(sum(accs[k]))
        return sh.cond_sampler(z[lo:hi], "0", "01234", m, sde1, x_init=x0[lo:hi], global_batch=16,
                               reduce_fn=reduce_fn, num_steps=1)
    sh.manual_seed(77)
    full = sh.cond_sampler(z, "0", "01234", m, sde1, x_init=x0, num_steps=1)
    one(0, 8, 0); one(8, 16, 0)
    lo = one(0, 8, 1); hi = one(8, 16, 1)
    assert rel_max(torch.cat([lo, hi]), full) < 1e-5


def test_dsm_loss_forward_backward_kernels_vs_oracle():
    sh = _sh()
    g = golden("dsm_loss.pt")
    batch, u, z = g["batch"], g["u"], g["z"]
    for c in g["cases"]:
        sde = _mk(c["kind"], c["a"], c["b"], c["N"])
        spec = so.SdeSpec(c["kind"], c["a"], c["b"], c["N"])
        w = torch.randn(5, 5, generator=torch.Generator().manual_seed(1)) * 0.3

        def toy(xx, tt, w=w):
            return torch.einsum("oc,bchw->bohw", w.to(xx.device), xx) * (1.0 + tt[:, None, None, None])

        wg = w.clone().cuda().requires_grad_(True)
        loss = sh.loss_fn(batch.cuda(), lambda xx, tt: toy(xx, tt, wg), sde, reduce_mean=c["reduce_mean"],
                          likelihood_weighting=c["likelihood_weighting"], u=u.cuda(), z=z.cuda())
        loss.backward()
        wr = w.clone().requires_grad_(True)
        ref = so.dsm_loss(spec, batch, lambda xx, tt: toy(xx, tt, wr), u, z, reduce_mean=c["reduce_mean"],
                          likelihood_weighting=c["likelihood_weighting"])
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item()), (c["kind"], loss.item(), ref.item())
        assert rel_l2(wg.grad, wr.grad) < 2e-5
