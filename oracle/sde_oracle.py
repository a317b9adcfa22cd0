"""ORACLE (test infrastructure, NOT product code).

CPU fp32 restatement of the reference's SDE / sampler / loss arithmetic
(/root/reference/sde_helper2.py and the inline conditional loop of
train_lat_celebhq_unet_cont2.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product package never does.

Every random draw of the reference (torch.randn_like / torch.rand) is an explicit argument
here so that the CUDA path and the oracle can be fed identical noise.

Pinned against the real reference modules by oracle/gen_golden.py (run in the build
container, where /root/reference exists); the resulting vectors live in tests/golden/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SdeSpec:
    """kind in {'vp','subvp','ve'}; (b0,b1) = (beta_min,beta_max) or (sigma_min,sigma_max); N steps; T = 1.
    sde_helper2.py:329-346, 384-400, 424-443."""
    kind: str
    b0: float
    b1: float
    N: int
    T: float = 1.0

    def alphas(self) -> torch.Tensor:
        # sde_helper2.py:342-343
        return 1.0 - torch.linspace(self.b0 / self.N, self.b1 / self.N, self.N)


def _bc(v: torch.Tensor) -> torch.Tensor:
    return v[:, None, None, None]


def sde_coeffs(s: SdeSpec, x: torch.Tensor, t: torch.Tensor):
    """(drift, diffusion[B]) of the forward SDE.  sde_helper2.py:352-356 (VP), 402-407 (subVP), 445-450 (VE)."""
    if s.kind in ("vp", "subvp"):
        beta_t = s.b0 + t * (s.b1 - s.b0)
        drift = -0.5 * _bc(beta_t) * x
        if s.kind == "vp":
            diffusion = torch.sqrt(beta_t)
        else:
            discount = 1.0 - torch.exp(-2 * s.b0 * t - (s.b1 - s.b0) * t ** 2)
            diffusion = torch.sqrt(beta_t * discount)
        return drift, diffusion
    sigma = s.b0 * (s.b1 / s.b0) ** t
    drift = torch.zeros_like(x)
    diffusion = sigma * torch.sqrt(torch.tensor(2 * (np.log(s.b1) - np.log(s.b0))))
    return drift, diffusion


def marginal_prob(s: SdeSpec, x: torch.Tensor, t: torch.Tensor):
    """(mean, std[B]) of p_t(x|x0).  sde_helper2.py:358-362, 409-413 (subVP std has NO sqrt), 452-455."""
    if s.kind in ("vp", "subvp"):
        lmc = -0.25 * t ** 2 * (s.b1 - s.b0) - 0.5 * t * s.b0
        mean = torch.exp(_bc(lmc)) * x
        std = torch.sqrt(1.0 - torch.exp(2.0 * lmc)) if s.kind == "vp" else 1 - torch.exp(2.0 * lmc)
        return mean, std
    return x, s.b0 * (s.b1 / s.b0) ** t


def reverse_sde(s: SdeSpec, x, t, score, probability_flow=False):
    """Reverse-time drift/diffusion given the score.  sde_helper2.py:277-317 (guidance branch off)."""
    drift, diffusion = sde_coeffs(s, x, t)
    drift = drift - _bc(diffusion) ** 2 * score * (0.5 if probability_flow else 1.0)
    if probability_flow:
        diffusion = torch.zeros_like(diffusion)
    return drift, diffusion


def em_predictor_step(s: SdeSpec, x, t, score, z, probability_flow=False):
    """Euler-Maruyama predictor.  sde_helper2.py:45-52.  `z` is the reference's randn_like(x) (drawn BEFORE the net call)."""
    dt = -1.0 / s.N
    drift, diffusion = reverse_sde(s, x, t, score, probability_flow)
    x_mean = x + drift * dt
    x_new = x_mean + _bc(diffusion) * np.sqrt(-dt) * z
    return x_new, x_mean


def discretize(s: SdeSpec, x, t):
    """(f, G[B]) of the discretised forward SDE, x_{i+1} = x_i + f + G z.
    subVP: the base class's Euler-Maruyama rule, sde_helper2.py:236-253 (f = drift/N, G = diffusion*sqrt(1/N));
    VP: DDPM rule, :373-381 (f = sqrt(alpha_i) x - x, G = sqrt(beta_i), i = (t(N-1)/T).long());
    VE: SMLD rule, :465-473 (f = 0, G = sqrt(sigma_i^2 - sigma_{i-1}^2), sigma_{-1} = 0)."""
    if s.kind == "subvp":
        dt = 1 / s.N
        drift, diffusion = sde_coeffs(s, x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt))
    timestep = (t * (s.N - 1) / s.T).long()
    if s.kind == "vp":
        betas = torch.linspace(s.b0 / s.N, s.b1 / s.N, s.N)
        beta = betas[timestep]
        alpha = (1.0 - betas)[timestep]
        return _bc(torch.sqrt(alpha)) * x - x, torch.sqrt(beta)
    sigmas = torch.exp(torch.linspace(np.log(s.b0), np.log(s.b1), s.N))
    sigma = sigmas[timestep]
    adjacent = torch.where(timestep == 0, torch.zeros_like(t), sigmas[timestep - 1])
    return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)


def rd_predictor_step(s: SdeSpec, x, t, score, z, probability_flow=False):
    """Reverse-diffusion predictor over RSDE.discretize (sde_helper2.py:319-324): rev_f = f - G^2 score [*0.5],
    rev_G = 0 if ODE else G; x_mean = x - rev_f; x = x_mean + rev_G z.  (The reference ships the discretisation
    without a caller; the update is the standard ancestral rule it was written for.)"""
    f, G = discretize(s, x, t)
    rev_f = f - _bc(G) ** 2 * score * (0.5 if probability_flow else 1.0)
    rev_G = torch.zeros_like(G) if probability_flow else G
    x_mean = x - rev_f
    return x_mean + _bc(rev_G) * z, x_mean


def corrector_alpha(s: SdeSpec, t: torch.Tensor) -> torch.Tensor:
    """sde_helper2.py:56-60: alpha looked up by truncating t*(N-1)/T to an integer index."""
    if s.kind in ("vp", "subvp"):
        timestep = (t * (s.N - 1) / s.T).long()
        return s.alphas()[timestep]
    return torch.ones_like(t)


def corrector_step(s: SdeSpec, x, t, grad, noise, target_snr):
    """One Langevin step.  sde_helper2.py:96-101.  `noise` is randn_like(x) drawn AFTER the net call.
    The step size couples the whole batch through the two .mean() calls."""
    alpha = corrector_alpha(s, t)
    grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
    step_size = (target_snr * noise_norm / grad_norm) ** 2 * 2 * alpha
    x_mean = x + _bc(step_size) * grad
    x_new = x_mean + _bc(torch.sqrt(step_size * 2)) * noise
    return x_new, x_mean


def impute_observed(s: SdeSpec, x, z_obs, obs_mask, t, noise_obs=True):
    """Overwrite observed modality channels with the (re-noised) clean latent.
    train_lat_celebhq_unet_cont2.py:293-303: noised = mean + std * z_obs where mean = exp(lmc) * z_obs,
    i.e. the clean latent is re-used as the 'noise' (no fresh Gaussian draw)."""
    out = x.clone()
    for m, on in enumerate(obs_mask):
        if not on:
            continue
        zm = z_obs[:, m:m + 1]
        if noise_obs:
            mean, std = marginal_prob(s, zm, t)
            out[:, m:m + 1] = mean + _bc(std) * zm
        else:
            out[:, m:m + 1] = zm
    return out


def timesteps(s: SdeSpec, eps: float) -> torch.Tensor:
    """sde_helper2.py:119 / train_lat_celebhq_unet_cont2.py:287."""
    return torch.linspace(s.T, eps, s.N)


def pc_sampler(s: SdeSpec, score_fn, x0, noise_pred, noise_corr, *, z_obs=None, obs_mask=None, eps=1e-3,
               noise_obs=True, pc=True, n_steps=1, target_snr=0.16, predictor_first=True, probability_flow=False,
               num_steps=None, return_trace=False, predictor="euler"):
    """N-step predictor-corrector sampler with observed-latent imputation.

    predictor_first=True  : train_lat_celebhq_unet_cont2.py:287-316 (calc_perf), train_poly_unet_cont.py:444-471
    predictor_first=False : sde_helper2.py:115-128 (uncond_sampler), train_lat_celebhq_unet_cont2.py:173-200
    x0          : initial stacked latent [B,M,D,D] (prior draw for the missing channels)
    noise_pred  : [steps,B,M,D,D] predictor noise; noise_corr: [steps,n_steps,B,M,D,D] corrector noise
    Returns the final latent: missing channels = last x_mean, observed channels = clean z_obs.
    """
    B = x0.shape[0]
    ts = timesteps(s, eps)
    steps = s.N if num_steps is None else num_steps
    rule = {"euler": em_predictor_step, "reverse_diffusion": rd_predictor_step}[predictor]
    x = x0.clone()
    x_mean = x0.clone()
    trace = []
    conditional = z_obs is not None and obs_mask is not None and any(obs_mask)
    for i in range(steps):
        vec_t = torch.ones(B) * ts[i]
        if conditional:
            x = impute_observed(s, x, z_obs, obs_mask, vec_t, noise_obs)

        def predictor(x):
            return rule(s, x, vec_t, score_fn(x, vec_t), noise_pred[i], probability_flow)

        def corrector(x):
            xm = x
            for k in range(n_steps):
                x, xm = corrector_step(s, x, vec_t, score_fn(x, vec_t), noise_corr[i, k], target_snr)
            return x, xm

        if predictor_first:
            x, x_mean = predictor(x)
            if pc:
                x, x_mean = corrector(x)
        else:
            if pc:
                x, x_mean = corrector(x)
            x, x_mean = predictor(x)
        if return_trace:
            trace.append(x.clone())
    out = x_mean.clone()
    if conditional:
        for m, on in enumerate(obs_mask):
            if on:
                out[:, m] = z_obs[:, m]
    return (out, trace) if return_trace else out


def likelihood_importance_cum_weight(t, beta_0, beta_1, eps=1e-5):
    """sde_helper2.py:129-134 (jax.numpy -> numpy)."""
    e1 = 0.5 * eps * (eps - 2) * beta_0 - 0.5 * eps ** 2 * beta_1
    e2 = 0.5 * t * (t - 2) * beta_0 - 0.5 * t ** 2 * beta_1
    term1 = np.where(np.abs(e1) <= 1e-3, -e1, 1.0 - np.exp(e1))
    term2 = np.where(np.abs(e2) <= 1e-3, -e2, 1.0 - np.exp(e2))
    return 0.5 * (-2 * np.log(term1) + 2 * np.log(term2) + beta_0 * (-2 * eps + eps ** 2 - (t - 2) * t)
                  + beta_1 * (-eps ** 2 + t ** 2))


def importance_sampled_t(s: SdeSpec, u01: torch.Tensor, eps=1e-5, steps=100) -> torch.Tensor:
    """sde_helper2.py:136-148: quantile = Uniform(0, Z).sample((B,)) (= u01 * Z with u01 = the underlying torch.rand
    draw, float32), then 100 bisection steps of the cumulative weight on [eps, T] in float32."""
    Z = likelihood_importance_cum_weight(s.T, s.b0, s.b1, eps)
    quantile = (0.0 + u01 * (float(Z) - 0.0)).numpy()
    lb = np.ones_like(quantile) * eps
    ub = np.ones_like(quantile) * s.T
    for _ in range(steps):
        mid = (lb + ub) / 2.0
        value = likelihood_importance_cum_weight(mid, s.b0, s.b1, eps=eps)
        lb = np.where(value <= quantile, mid, lb)
        ub = np.where(value <= quantile, ub, mid)
    return torch.tensor(np.array((lb + ub) / 2.0))


def dsm_loss(s: SdeSpec, batch, score_fn, u, z, *, reduce_mean=True, likelihood_weighting=False, eps=1e-5, t_is=None):
    """Denoising score matching loss.  sde_helper2.py:152-186.
    u = torch.rand(B) and z = torch.randn_like(batch) are the reference's two draws, in that order.
    t_is: the importance-sampled times of the `likelihood_weighting and im_sample` branch (:164-165, 177-179;
    `importance_sampled_t`): t = t_is, loss = red((score * std + z)^2) without the g^2 weight."""
    if t_is is not None:
        t = t_is.to(batch.dtype)
        mean, std = marginal_prob(s, batch, t)
        score = score_fn(mean + _bc(std) * z, t)
        losses = torch.square(score * _bc(std) + z).reshape(batch.shape[0], -1)
        return torch.mean(torch.mean(losses, dim=-1) if reduce_mean else 0.5 * torch.sum(losses, dim=-1))
    t = u * (s.T - eps) + eps
    mean, std = marginal_prob(s, batch, t)
    perturbed = mean + _bc(std) * z
    score = score_fn(perturbed, t)
    if reduce_mean:
        red = lambda a: torch.mean(a, dim=-1)
    else:
        red = lambda a: 0.5 * torch.sum(a, dim=-1)
    if not likelihood_weighting:
        losses = torch.square(score * _bc(std) + z)
        losses = red(losses.reshape(losses.shape[0], -1))
    else:
        g2 = sde_coeffs(s, torch.zeros_like(batch), t)[1] ** 2
        losses = torch.square(score + z / _bc(std))
        losses = red(losses.reshape(losses.shape[0], -1)) * g2
    return torch.mean(losses)


def prior_logp(s: SdeSpec, z: torch.Tensor) -> torch.Tensor:
    """sde_helper2.py:367-371, 418-421, 460-463."""
    n = float(np.prod(z.shape[1:]))
    if s.kind == "ve":
        return -n / 2.0 * math.log(2 * math.pi * s.b1 ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * s.b1 ** 2)
    return -n / 2.0 * math.log(2 * math.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.0


def annealed_langevin(model, z_all, obs_mask, er, c, sigmas, n_comp, noise, num_levels=None):
    """Legacy annealed-Langevin evaluator, eval_lat_celeba_hq_all.py:258-275, on the stacked latent [B,M,D,D]
    (the reference keeps a dict of [B,size_z] latents and re-stacks it before every net call: same values).
    er, c: per-channel lists; sigmas: float64 numpy levels; noise: [levels, n_comp, B, M, D, D] (the reference draws one
    randn_like per missing modality: channel m of noise[s, i] stands for that draw).  Restated from an inline script
    loop (not an importable function): pinned by reading, not by running the reference."""
    x = z_all.clone()
    sig = torch.tensor(sigmas)                                           # float64, :222
    B = x.shape[0]
    levels = len(sigmas) if num_levels is None else num_levels
    for s_in in range(levels):
        sigma_index = torch.tensor([s_in] * B)
        cur = sig[sigma_index].float()
        for i in range(n_comp):
            sm_out = model(x, sigma_index) / cur.view(B, 1, 1, 1)
            new = x.clone()
            for m, on in enumerate(obs_mask):
                if not on:
                    alpha = er[m] * (sig[s_in] ** 2) / (sig[-1] ** 2)
                    new[:, m] = x[:, m] + (alpha * sm_out[:, m]) + c[m] * (torch.sqrt(2 * alpha) * noise[s_in, i][:, m])
            x = new.float()
    return x


def langevin_refine(sm_model, z_all, obs_mask, n_comp, lr1, lr2, schedule, noise):
    """Fixed-step evaluator, fid_upd10.py:279-290 (inline script loop, restated).  noise: [draws, B, M, D, D]."""
    x = z_all.clone()
    B = x.shape[0]
    k = 0
    for i in range(n_comp):
        sm_out = sm_model(x.view(B, -1)).view_as(x)
        steps = [lr1] if not schedule else [lr1 * ((i + 1) / n_comp)] + ([1 * ((i + 1) / n_comp)] if i == n_comp - 1 else [])
        for a in steps:
            new = x.clone()
            for m, on in enumerate(obs_mask):
                if not on:
                    new[:, m] = x[:, m] + (a * sm_out[:, m]) + lr2 * noise[k][:, m]
            x = new
            k += 1
    return x
