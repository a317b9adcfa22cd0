"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (/root/reference) on CPU,
and check the oracle restatement (oracle/sde_oracle.py, oracle/unet_oracle.py) against them.

Run in the build container only (the GPU box has no /root/reference):
    python -m oracle.gen_golden
The reference imports `jax.numpy` (absent here) for one unused-by-default helper; a numpy shim stands
in for it (sde_helper2.py:5,131-150 only use where/abs/exp/log/ones_like).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

from . import sde_oracle as so
from . import unet_oracle as uo
from .det_weights import fill_state_dict

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    if "jax" not in sys.modules:
        jax = types.ModuleType("jax")
        jax.numpy = np
        sys.modules["jax"] = jax
        sys.modules["jax.numpy"] = np
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import sde_helper2
    import unet_model
    import unet_openai
    return sde_helper2, unet_model, unet_openai


class NoiseFeed:
    """Replays pre-drawn tensors through torch.randn_like / torch.rand / torch.randn, in call order."""

    def __init__(self, normals=(), uniforms=()):
        self.normals = list(normals)
        self.uniforms = list(uniforms)

    @contextlib.contextmanager
    def patched(self):
        o_rl, o_r, o_rn = torch.randn_like, torch.rand, torch.randn

        def randn_like(x, **kw):
            t = self.normals.pop(0)
            assert t.shape == x.shape, (t.shape, x.shape)
            return t.clone()

        def rand(*shape, **kw):
            t = self.uniforms.pop(0)
            return t.clone()

        def randn(*shape, **kw):
            t = self.normals.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        torch.randn_like, torch.rand, torch.randn = randn_like, rand, randn
        try:
            yield
        finally:
            torch.randn_like, torch.rand, torch.randn = o_rl, o_r, o_rn


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main():
    torch.set_num_threads(8)
    sh, um, uoa = import_reference()
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(20240607)
    report = {}

    # ------------------------------------------------------------------ score nets
    nets = {}
    for name, kw, xshape in [
        ("unet_poly", dict(dim=32, channels=5, dim_mults=(1, 2, 2, 2)), (3, 5, 8, 8)),
        ("unet_cel", dict(dim=32, channels=3, dim_mults=(1, 2, 2, 2, 2)), (2, 3, 16, 16)),
    ]:
        ref = um.Unet(**kw).eval()
        shapes = shapes_of(ref)
        sd = fill_state_dict(shapes)
        ref.load_state_dict(sd)
        x = torch.randn(xshape, generator=g)
        t = torch.rand(xshape[0], generator=g) * 0.999 + 1e-3
        with torch.no_grad():
            y = ref(x, t)
            y_or = uo.unet_forward(sd, x, t, dim=kw["dim"], dim_mults=kw["dim_mults"])
        report[name + "/oracle_vs_ref"] = rel(y_or, y)
        assert rel(y_or, y) < 1e-5, report
        torch.save({"kwargs": kw, "shapes": shapes, "x": x, "t": t, "y": y}, os.path.join(OUT, name + ".pt"))
        nets[name] = (ref, sd, kw)

    kw = dict(in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1, attention_resolutions=(2,),
              dropout=0.1, channel_mult=(1, 2, 2), num_heads=2, use_z=True, z_dim=16)
    ref = uoa.UNetModel(**kw).eval()
    shapes = shapes_of(ref)
    sd = fill_state_dict(shapes)
    ref.load_state_dict(sd)
    x = torch.randn(2, 3, 8, 8, generator=g)
    t = torch.rand(2, generator=g) * 0.999 + 1e-3
    z = torch.randn(2, 16, generator=g)
    with torch.no_grad():
        y = ref(x, t, z=z)
        y0 = ref(x, t)
        y_or = uo.unet_openai_forward(sd, x, t, model_channels=32, num_res_blocks=1, attention_resolutions=(2,),
                                      channel_mult=(1, 2, 2), num_heads=2, z=z)
        y0_or = uo.unet_openai_forward(sd, x, t, model_channels=32, num_res_blocks=1, attention_resolutions=(2,),
                                       channel_mult=(1, 2, 2), num_heads=2)
    report["unet_openai/oracle_vs_ref"] = max(rel(y_or, y), rel(y0_or, y0))
    assert report["unet_openai/oracle_vs_ref"] < 1e-5, report
    torch.save({"kwargs": kw, "shapes": shapes, "x": x, "t": t, "z": z, "y": y, "y_noz": y0},
               os.path.join(OUT, "unet_openai.pt"))

    # ------------------------------------------------------------------ SDE objects
    sde_cases = []
    for kind, cls, a, b, N in [("vp", sh.VPSDE, 0.1, 20.0, 1000), ("vp", sh.VPSDE, 1.0, 5.0, 100),
                               ("subvp", sh.subVPSDE, 0.1, 20.0, 50), ("ve", sh.VESDE, 0.01, 50.0, 30)]:
        sde = cls(a, b, N)
        spec = so.SdeSpec(kind, a, b, N)
        x = torch.randn(4, 3, 4, 4, generator=g)
        t = torch.rand(4, generator=g) * 0.999 + 1e-3
        drift, diff = sde.sde(x, t)
        mean, std = sde.marginal_prob(x, t)
        logp = sde.prior_logp(x)
        d_or, g_or = so.sde_coeffs(spec, x, t)
        m_or, s_or = so.marginal_prob(spec, x, t)
        assert torch.equal(d_or, drift) and torch.allclose(g_or, diff, rtol=1e-6), kind
        assert torch.equal(m_or, mean) and torch.equal(s_or, std), kind
        assert torch.allclose(so.prior_logp(spec, x), logp, rtol=1e-6), kind
        case = {"kind": kind, "a": a, "b": b, "N": N, "x": x, "t": t, "drift": drift, "diffusion": diff,
                "mean": mean, "std": std, "prior_logp": logp}
        if kind == "vp":
            f, G = sde.discretize(x, t)
            case.update({"disc_f": f, "disc_G": G, "alphas": sde.alphas.clone(),
                         "sqrt_1m_alphas_cumprod": sde.sqrt_1m_alphas_cumprod.clone()})
        if kind == "subvp":
            f, G = sde.discretize(x, t)
            case.update({"disc_f": f, "disc_G": G})
        sde_cases.append(case)
    torch.save(sde_cases, os.path.join(OUT, "sde_objects.pt"))

    # ------------------------------------------------------------------ sampler steps (VP, Poly-like net)
    ref, sd, kw = nets["unet_poly"]
    score_or = lambda x, t: uo.unet_forward(sd, x, t, dim=kw["dim"], dim_mults=kw["dim_mults"])
    sde = sh.VPSDE(1.0, 5.0, 10)
    spec = so.SdeSpec("vp", 1.0, 5.0, 10)
    B = 4
    x = torch.randn(B, 5, 8, 8, generator=g)
    vec_t = torch.ones(B) * 0.77
    z1 = torch.randn(B, 5, 8, 8, generator=g)
    z2 = torch.randn(B, 5, 8, 8, generator=g)
    with torch.no_grad():
        with NoiseFeed([z1]).patched():
            xp, xp_mean = sh.em_predictor(x, vec_t, ref, sde)
        # NOTE: sh.em_predictor(..., probability_flow=True) raises TypeError in the reference
        # (sde_helper2.py:316 sets diffusion = 0. and :51 then subscripts it), so the ODE predictor is unpinned.
        try:
            sh.em_predictor(x, vec_t, ref, sde, probability_flow=True)
            ode_raises = False
        except TypeError:
            ode_raises = True
        assert ode_raises
        with NoiseFeed([z2]).patched():
            xc, xc_mean = sh.corrector(x, vec_t, ref, sde, 1, 0.16)
        score = ref(x, vec_t)
    a, b_ = so.em_predictor_step(spec, x, vec_t, score, z1)
    assert rel(a, xp) < 1e-6 and rel(b_, xp_mean) < 1e-6
    a, b_ = so.corrector_step(spec, x, vec_t, score, z2, 0.16)
    assert rel(a, xc) < 1e-6 and rel(b_, xc_mean) < 1e-6
    steps = {"sde": ("vp", 1.0, 5.0, 10), "x": x, "t": vec_t, "z_pred": z1, "z_corr": z2, "score": score,
             "pred_x": xp, "pred_mean": xp_mean, "ode_raises_in_reference": ode_raises, "corr_x": xc,
             "corr_mean": xc_mean, "target_snr": 0.16}

    # conditional loop, restated verbatim from train_lat_celebhq_unet_cont2.py:287-316 / 173-200 using the
    # REFERENCE em_predictor / corrector / marginal_prob (the scripts themselves are not importable)
    def ref_cond_loop(z0, given, mods, nsteps, predictor_first, noise_obs, npred, ncorr):
        zz = {m: z0[:, i].reshape(B, 64).clone() for i, m in enumerate(mods)}
        ts = torch.linspace(sde.T, 1e-3, sde.N)
        noised = {}
        feed = []
        for i in range(nsteps):
            feed += [npred[i], ncorr[i, 0]] if predictor_first else [ncorr[i, 0], npred[i]]
        nf = NoiseFeed(feed)
        with torch.no_grad(), nf.patched():
            for i in range(nsteps):
                vt = torch.ones(B) * ts[i]
                for m in mods:
                    if noise_obs and m in given:
                        mean, std = sde.marginal_prob(zz[m].view(-1, 1, 8, 8), vt)
                        noised[m] = (mean + std[:, None, None, None] * zz[m].view(-1, 1, 8, 8)).view(-1, 64)
                    else:
                        noised[m] = zz[m]
                z_upd = torch.cat([noised[m].unsqueeze(1) for m in mods], dim=1).view(-1, len(mods), 8, 8)
                if predictor_first:
                    z_upd, z_mean = sh.em_predictor(z_upd, vt, ref, sde, given=given, all_mods=mods)
                    z_upd, z_mean = sh.corrector(z_upd, vt, ref, sde, 1, 0.16, given=given, all_mods=mods)
                else:
                    z_upd, z_mean = sh.corrector(z_upd, vt, ref, sde, 1, 0.16, given=given, all_mods=mods)
                    z_upd, z_mean = sh.em_predictor(z_upd, vt, ref, sde, given=given, all_mods=mods)
                for ind, m in enumerate(mods):
                    if m not in given:
                        zz[m] = z_upd[:, ind].reshape(B, 64)
            for ind, m in enumerate(mods):
                if m not in given:
                    zz[m] = z_mean[:, ind].reshape(B, 64)
        return torch.stack([zz[m] for m in mods], dim=1).view(B, len(mods), 8, 8)

    mods = "01234"
    nsteps = 3
    z0 = torch.randn(B, 5, 8, 8, generator=g)
    npred = torch.randn(nsteps, B, 5, 8, 8, generator=g)
    ncorr = torch.randn(nsteps, 1, B, 5, 8, 8, generator=g)
    loops = []
    for given, pf, nobs in [("0", True, True), ("13", True, True), ("0", False, True), ("24", True, False),
                            ("", True, True)]:
        out = ref_cond_loop(z0, given, mods, nsteps, pf, nobs, npred, ncorr)
        mask = [m in given for m in mods]
        with torch.no_grad():
            out_or = so.pc_sampler(spec, score_or, z0, npred, ncorr, z_obs=z0, obs_mask=mask, noise_obs=nobs,
                                   predictor_first=pf, num_steps=nsteps)
        r = rel(out_or, out)
        report[f"cond_loop/{given or 'uncond'}/{'pc' if pf else 'cp'}"] = r
        assert r < 1e-5, report
        loops.append({"given": given, "predictor_first": pf, "noise_obs": nobs, "out": out})
    steps.update({"loop_z0": z0, "loop_npred": npred, "loop_ncorr": ncorr, "loop_steps": nsteps, "loops": loops,
                  "mods": mods})

    # library uncond_sampler (corrector -> predictor), full N = 10 steps
    prior = torch.randn(B, 5, 8, 8, generator=g)
    up = torch.randn(10, B, 5, 8, 8, generator=g)
    uc = torch.randn(10, 1, B, 5, 8, 8, generator=g)
    feed = [prior]
    for i in range(10):
        feed += [uc[i, 0], up[i]]
    with NoiseFeed(feed).patched():
        u_out = sh.uncond_sampler((B, 5, 8, 8), ref, torch.device("cpu"), sde, pc=True)
    with torch.no_grad():
        u_or = so.pc_sampler(spec, score_or, prior, up, uc, predictor_first=False)
    report["uncond_sampler"] = rel(u_or, u_out)
    assert report["uncond_sampler"] < 1e-4, report
    steps.update({"uncond_prior": prior, "uncond_npred": up, "uncond_ncorr": uc, "uncond_out": u_out})
    torch.save(steps, os.path.join(OUT, "sampler_steps.pt"))

    # ------------------------------------------------------------------ DSM loss (+ gradients)
    batch = torch.randn(B, 5, 8, 8, generator=g)
    u = torch.rand(B, generator=g)
    zn = torch.randn(B, 5, 8, 8, generator=g)
    losses = []
    grad_keys = ["final_conv.1.weight", "downs.0.0.net.1.weight", "time_mlp.1.weight", "mid_attn.fn.fn.to_qkv.weight",
                 "ups.2.3.weight", "downs.1.2.fn.fn.to_out.1.weight", "init_conv.bias"]
    for kind, cls, a, b_, rm, lw in [("vp", sh.VPSDE, 1.0, 5.0, True, False), ("vp", sh.VPSDE, 0.1, 20.0, True, True),
                                     ("vp", sh.VPSDE, 0.1, 20.0, False, True),
                                     ("subvp", sh.subVPSDE, 0.1, 20.0, True, False),
                                     ("ve", sh.VESDE, 0.01, 50.0, False, False)]:
        sde_l = cls(a, b_, 100)
        spec_l = so.SdeSpec(kind, a, b_, 100)
        ref.zero_grad()
        with NoiseFeed([zn], [u]).patched():
            loss = sh.loss_fn(batch, ref, sde_l, reduce_mean=rm, likelihood_weighting=lw)
        loss.backward()
        params = dict(ref.named_parameters())
        grads = {k: {"norm": params[k].grad.norm().clone(), "head": params[k].grad.flatten()[:512].clone()}
                 for k in grad_keys}
        gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in ref.parameters()))
        with torch.no_grad():
            l_or = so.dsm_loss(spec_l, batch, score_or, u, zn, reduce_mean=rm, likelihood_weighting=lw)
        r = abs(l_or.item() - loss.item()) / abs(loss.item())
        report[f"loss/{kind}/rm{int(rm)}lw{int(lw)}"] = r
        assert r < 1e-5, report
        losses.append({"kind": kind, "a": a, "b": b_, "N": 100, "reduce_mean": rm, "likelihood_weighting": lw,
                       "loss": loss.detach().clone(), "grads": grads, "grad_norm": gnorm.detach().clone()})
    torch.save({"batch": batch, "u": u, "z": zn, "cases": losses}, os.path.join(OUT, "dsm_loss.pt"))

    for k, v in report.items():
        print(f"{k:40s} oracle-vs-reference rel err {v:.3e}")
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT)) if f.endswith('.pt')}
    print("fixtures:", sizes)


if __name__ == "__main__":
    main()
