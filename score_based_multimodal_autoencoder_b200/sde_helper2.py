"""Drop-in for the reference's `sde_helper2.py`: SDE objects, predictor / corrector, samplers, DSM loss.

Same names, positional order and defaults as the reference:
    VPSDE / subVPSDE / VESDE, em_predictor, corrector, uncond_sampler, loss_fn
plus the additive `cond_sampler`, which lifts the conditional predictor-corrector loop that the
reference copy-pastes inline (train_lat_celebhq_unet_cont2.py:263-316 and five other sites) into a
library entry point.

The step arithmetic runs in the fused kernels of csrc/sampler.cu through the C ABI.  The small SDE-object
methods (`sde()`, `marginal_prob()`, ...) are kept as host-side tensor expressions for API completeness
(the reference's scripts call them directly); the samplers themselves never go through them.

Noise.  The reference draws `torch.randn_like` (sde_helper2.py:47, 96, 168).  Three sources are supported:
  * `noise=` / `u=`, `z=` keyword arguments: injected tensors (parity tests);
  * rng="torch":  draw with torch's generator in the reference's call order, so the same `torch.manual_seed`
    reproduces the reference's own CUDA stream on the same device (default for the standalone step functions);
  * rng="philox": in-kernel Philox4x32-10 keyed by global element index (default for the N-step samplers;
    nothing but x and the score touches HBM, and a batch shard draws what the full batch would).
"""
from __future__ import annotations

import abc
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_KIND = {"vp": L.SDE_VP, "subvp": L.SDE_SUBVP, "ve": L.SDE_VE}


def _bc(v):
    return v[:, None, None, None]


# ======================================================================================= SDE objects
class SDE(abc.ABC):
    """sde_helper2.py:191-326."""

    _kind = None

    def __init__(self, N):
        super().__init__()
        self.N = N
        self._dev_tables = {}

    @property
    @abc.abstractmethod
    def T(self):
        ...

    @abc.abstractmethod
    def sde(self, x, t):
        ...

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        ...

    @abc.abstractmethod
    def prior_sampling(self, shape):
        ...

    @abc.abstractmethod
    def prior_logp(self, z):
        ...

    def discretize(self, x, t):
        """Euler-Maruyama discretisation, sde_helper2.py:236-253."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE/ODE object (sde_helper2.py:255-326): `.sde(x, t, ...)`, `.discretize(x, t)`."""
        return _ReverseSDE(self, score_fn, probability_flow)

    # ---- C-ABI view
    def _c(self) -> L.SdeC:
        return L.SdeC(_KIND[self._kind], float(self.beta_0), float(self.beta_1), int(self.N), float(self.T))

    def _alphas_on(self, device):
        """Device-resident copy of `alphas` (the reference re-uploads the table on every corrector call,
        sde_helper2.py:58)."""
        if not hasattr(self, "alphas"):
            return None
        key = str(device)
        tab = self._dev_tables.get(key)
        if tab is None:
            tab = self.alphas.to(device=device, dtype=torch.float32).contiguous()
            self._dev_tables[key] = tab
        return tab


class _ReverseSDE:
    def __init__(self, fwd: SDE, score_fn, probability_flow):
        self._fwd, self._score_fn = fwd, score_fn
        self.N, self.probability_flow = fwd.N, probability_flow

    @property
    def T(self):
        return self._fwd.T

    def sde(self, x, t, cl_g=None, cl_s=None, target=None, given=None, all_mods=None):
        _reject_guidance(cl_g, given)
        drift, diffusion = self._fwd.sde(x, t)
        score = self._score_fn(x, t)
        drift = drift - _bc(diffusion) ** 2 * score * (0.5 if self.probability_flow else 1.0)
        # the reference returns the float 0. here (sde_helper2.py:316), which its own em_predictor cannot index;
        # a zero vector keeps the contract usable
        diffusion = torch.zeros_like(diffusion) if self.probability_flow else diffusion
        return drift, diffusion

    def discretize(self, x, t):
        f, G = self._fwd.discretize(x, t)
        rev_f = f - _bc(G) ** 2 * self._score_fn(x, t) * (0.5 if self.probability_flow else 1.0)
        rev_G = torch.zeros_like(G) if self.probability_flow else G
        return rev_f, rev_G


class VPSDE(SDE):
    """sde_helper2.py:329-381."""
    _kind = "vp"

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0 = beta_min
        self.beta_1 = beta_max
        self.N = N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * _bc(beta_t) * x, torch.sqrt(beta_t)

    def marginal_prob(self, x, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(_bc(lmc)) * x, torch.sqrt(1.0 - torch.exp(2.0 * lmc))

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.0

    def discretize(self, x, t):
        """DDPM discretisation, sde_helper2.py:373-381."""
        timestep = (t * (self.N - 1) / self.T).long()
        beta = self.discrete_betas.to(x.device)[timestep]
        alpha = self.alphas.to(x.device)[timestep]
        return _bc(torch.sqrt(alpha)) * x - x, torch.sqrt(beta)


class subVPSDE(SDE):
    """sde_helper2.py:384-421."""
    _kind = "subvp"

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0 = beta_min
        self.beta_1 = beta_max
        self.N = N
        # the reference's corrector reads sde.alphas for subVPSDE too (sde_helper2.py:56-58) although the class
        # never defines it (AttributeError upstream); the VP table is the evident intent
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1.0 - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * _bc(beta_t) * x, torch.sqrt(beta_t * discount)

    def marginal_prob(self, x, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return _bc(torch.exp(lmc)) * x, 1 - torch.exp(2.0 * lmc)

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.0


class VESDE(SDE):
    """sde_helper2.py:424-473."""
    _kind = "ve"

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
        super().__init__(N)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.beta_0 = self.sigma_min
        self.beta_1 = self.sigma_max
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(self.sigma_min), np.log(self.sigma_max), N))
        self.N = N

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        diffusion = sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)),
                                                    device=t.device))
        return torch.zeros_like(x), diffusion

    def marginal_prob(self, x, t):
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def prior_sampling(self, shape):
        return torch.randn(*shape) * self.sigma_max

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return (-n / 2.0 * np.log(2 * np.pi * self.sigma_max ** 2)
                - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2))

    def discretize(self, x, t):
        """SMLD discretisation, sde_helper2.py:465-473 (table moved to t's device first; the reference indexes a CPU
        table with a device index and fails on CUDA)."""
        timestep = (t * (self.N - 1) / self.T).long()
        sig = self.discrete_sigmas.to(t.device)
        sigma = sig[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(t), sig[timestep - 1])
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)


# ======================================================================================= helpers
def _reject_guidance(cl_g, given):
    if cl_g is not None and given is not None and given:
        raise NotImplementedError(
            "classifier/EBM guidance (sde_helper2.py:65-94, 283-312) needs the ClwithTime2/3 energy nets, which the "
            "reference repository does not contain; pass cl_g=None (the shipped default, --use-clg 0)")


def _latent_shape(x) -> L.LatentShape:
    if x.dim() != 4:
        raise ValueError("latent must be [B, M, D, D]")
    b, m, d1, d2 = x.shape
    return L.LatentShape(b, m, d1 * d2)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.SbmError("the B200 sampler path needs CUDA tensors (no CPU fallback)")


def _f32c(t):
    return t.detach().contiguous().float()


class _RngState:
    """Host-side bookkeeping of the in-kernel Philox stream (seed + monotonically increasing draw id)."""

    def __init__(self):
        self.seed = 0x5B3AE_B200
        self.draw = 0
        self.sample_offset = 0

    def next(self, draw_dev=None) -> L.Rng:
        r = L.Rng(self.seed, self.draw, self.sample_offset, draw_dev.data_ptr() if draw_dev is not None else None)
        self.draw += 1
        return r


_rng = _RngState()


def manual_seed(seed: int, sample_offset: int = 0) -> None:
    """Seed the in-kernel Philox generator.  `sample_offset` = index of this rank's first sample in the global
    batch, so that sharded sampling draws exactly what the unsharded batch would."""
    _rng.seed, _rng.draw, _rng.sample_offset = int(seed) & (2 ** 64 - 1), 0, int(sample_offset)


def _impute_struct(z_obs, obs_mask, noise_obs, t_next, t_next_dev=None):
    if z_obs is None or not obs_mask:
        return None
    return L.Impute(z_obs.data_ptr(), int(obs_mask), 1 if noise_obs else 0, float(t_next),
                    t_next_dev.data_ptr() if t_next_dev is not None else None)


def _predictor_kernel(sde, x, score, t, *, noise=None, rng=None, probability_flow=False, impute=None, want_mean=True,
                      out=None):
    ls = _latent_shape(x)
    x_new = out if out is not None else torch.empty_like(x)
    x_mean = torch.empty_like(x) if want_mean else None
    sc = sde._c()
    L.check(L.lib().sbm_predictor_step(C.byref(ls), C.byref(sc), L.ptr(x), L.ptr(score), L.ptr(t), L.ptr(noise),
                                       L.ptr(x_new), L.ptr(x_mean), C.c_int32(1 if probability_flow else 0),
                                       C.byref(rng) if rng is not None else None,
                                       C.byref(impute) if impute is not None else None, L.stream_ptr()),
            "sbm_predictor_step")
    return x_new, x_mean


def _corrector_kernels(sde, x, grad, t, target_snr, *, noise=None, rng=None, impute=None, want_mean=True,
                       global_batch=None, reduce_fn=None, acc=None, out=None):
    """norms kernel -> (optional cross-rank sum of the two batch norms) -> update kernel.  `acc` is a zeroed buffer of
    3 doubles (2 sums + a completion ticket); the update kernel re-zeroes it, so a reused `acc` never needs a memset."""
    ls = _latent_shape(x)
    if acc is None:
        acc = torch.zeros(3, dtype=torch.float64, device=x.device)
    elif acc.numel() < 3:
        raise L.SbmError("corrector accumulator needs 3 doubles (2 sums + ticket)")
    rp = C.byref(rng) if rng is not None else None
    L.check(L.lib().sbm_corrector_norms(C.byref(ls), L.ptr(grad), L.ptr(noise), rp, L.ptr(acc), L.stream_ptr()),
            "sbm_corrector_norms")
    if reduce_fn is not None:  # multi-GPU exact mode: sum the two batch norms over ranks
        reduce_fn(acc[:2])
    x_new = out if out is not None else torch.empty_like(x)
    x_mean = torch.empty_like(x) if want_mean else None
    sc = sde._c()
    alphas = sde._alphas_on(x.device) if isinstance(sde, (VPSDE, subVPSDE)) else None
    L.check(L.lib().sbm_corrector_update(C.byref(ls), C.byref(sc), L.ptr(x), L.ptr(grad), L.ptr(t), L.ptr(noise),
                                         L.ptr(acc), L.ptr(alphas), L.ptr(x_new), L.ptr(x_mean),
                                         C.c_float(target_snr), C.c_int64(global_batch or x.shape[0]), rp,
                                         C.byref(impute) if impute is not None else None, C.c_int32(1),
                                         L.stream_ptr()),
            "sbm_corrector_update")
    return x_new, x_mean


# ======================================================================================= step functions
def _call_score(score_fn, x, t, z_cond):
    """score_fn(x, t) of the reference; with a conditioning code the z-conditioned call of
    train_lat_celebhq_unet_cont2_cond.py:123,225-226,307-309 (`score_fn(x, t, z=z_cond)`, UNetModel(use_z=True))."""
    return score_fn(x, t) if z_cond is None else score_fn(x, t, z=z_cond)


def em_predictor(x, t, score_fn, sde, probability_flow=False, cl_g=None, cl_s=None, target=None, given=None,
                 all_mods=None, z_cond=None, *, noise=None, rng="torch"):
    """Euler-Maruyama reverse-SDE predictor step (sde_helper2.py:45-52) -> (x, x_mean).
    One fused kernel after the score-net call; the noise is drawn BEFORE the net call like the reference."""
    _reject_guidance(cl_g, given)
    _need_cuda(x, t)
    x, t = _f32c(x), _f32c(t)
    r = None
    if noise is None and not probability_flow:
        if rng == "torch":
            noise = torch.randn_like(x)
        else:
            r = _rng.next()
    score = _f32c(_call_score(score_fn, x, t, z_cond))
    return _predictor_kernel(sde, x, score, t, noise=noise, rng=r, probability_flow=probability_flow)


def rd_predictor(x, t, score_fn, sde, probability_flow=False, z_cond=None, *, noise=None):
    """Reverse-diffusion (ancestral) predictor step -> (x, x_mean):
        (rev_f, rev_G) = sde.reverse(score_fn, probability_flow).discretize(x, t);  x_mean = x - rev_f;
        x = x_mean + rev_G * z.
    The reference ships this discretisation (sde_helper2.py:236-253, 319-324, 373-381) but no caller;
    this is the API-complete thin wrapper over the SDE classes' tensor expressions -- a handful of elementwise torch
    ops on the caller's device next to the score-net call, not a fused kernel and not on the measured path."""
    fn = score_fn if z_cond is None else (lambda a, b: _call_score(score_fn, a, b, z_cond))
    z = torch.randn_like(x) if noise is None else noise     # drawn before the net call, like em_predictor (:47)
    rev_f, rev_G = sde.reverse(fn, probability_flow).discretize(x, t)
    x_mean = x - rev_f
    return x_mean + _bc(rev_G) * z, x_mean


def corrector(x, t, score_fn, sde, n_steps, target_snr, cl_g=None, cl_s=None, target=None, given=None, all_mods=None,
              z_cond=None, *, noise=None, rng="torch", global_batch=None, reduce_fn=None):
    """Langevin corrector (sde_helper2.py:54-106) -> (x, x_mean).  Two fused kernels per Langevin step
    (batch-coupled norms, then the update); the noise is drawn AFTER the net call like the reference.
    `noise`: optional [n_steps, B, M, D, D] injected noise."""
    _reject_guidance(cl_g, given)
    _need_cuda(x, t)
    x, t = _f32c(x), _f32c(t)
    x_mean = x
    for i in range(n_steps):
        grad = _f32c(_call_score(score_fn, x, t, z_cond))
        nz, r = None, None
        if noise is not None:
            nz = noise[i] if noise.dim() == 5 else noise
        elif rng == "torch":
            nz = torch.randn_like(x)
        else:
            r = _rng.next()
        x, x_mean = _corrector_kernels(sde, x, grad, t, target_snr, noise=nz, rng=r, global_batch=global_batch,
                                       reduce_fn=reduce_fn)
    return x, x_mean


# ======================================================================================= N-step samplers
def _obs_mask_from(given, all_mods) -> int:
    mask = 0
    if given:
        for i, m in enumerate(all_mods):
            if m in given:
                mask |= 1 << i
    return mask


@torch.no_grad()
def pc_sampler(x0, model, sde, *, z_obs=None, obs_mask=0, eps=1e-3, noise_obs=True, pc=True, n_steps=1,
               target_snr=0.16, predictor_first=True, probability_flow=False, noise_pred=None, noise_corr=None,
               num_steps=None, global_batch=None, reduce_fn=None, use_graph=False, return_state=False,
               rng="philox", z_cond=None):
    """N-step predictor-corrector sampler over a stacked latent [B,M,D,D] with observed-modality imputation.

    Semantics = the reference's inline loop (train_lat_celebhq_unet_cont2.py:287-316 for predictor_first=True,
    :173-200 / sde_helper2.py:121-128 for predictor_first=False).  Per step: 2 score-net forwards + 3 fused
    elementwise kernels (predictor, corrector norms, corrector update); the imputation of the observed channels
    for step i+1 is an epilogue of the last kernel of step i.  Returns the last x_mean with the observed
    channels set to the clean latents.
    """
    _need_cuda(x0)
    dev = x0.device
    x = _f32c(x0).clone()
    B = x.shape[0]
    N = sde.N
    steps = N if num_steps is None else num_steps
    ts = torch.linspace(sde.T, eps, N, device=dev)
    ts_host = ts.tolist()
    conditional = z_obs is not None and obs_mask != 0
    if conditional:
        z_obs = _f32c(z_obs)
        ls = _latent_shape(x)
        sc = sde._c()
        im0 = _impute_struct(z_obs, obs_mask, noise_obs, ts_host[0])
        L.check(L.lib().sbm_impute_observed(C.byref(ls), C.byref(sc), L.ptr(x), L.ptr(x), C.byref(im0),
                                            L.stream_ptr()), "sbm_impute_observed")
    inject = noise_pred is not None
    torch_rng = (rng == "torch") and not inject
    t_vec = torch.empty(B, device=dev, dtype=torch.float32)
    acc = torch.zeros(3, dtype=torch.float64, device=dev)
    x_mean = x

    def one_step(i, x, *, last, graph_state=None):
        """Runs step i; `graph_state` = (t_next_dev, draw_dev) when the step is being captured for replay."""
        t_next_dev, draw_dev = graph_state if graph_state is not None else (None, None)
        if graph_state is None:
            t_vec.fill_(ts_host[i])
        im = None
        if conditional and not last:
            im = _impute_struct(z_obs, obs_mask, noise_obs, ts_host[min(i + 1, N - 1)], t_next_dev)

        def predictor(x, impute, want_mean):
            nz = noise_pred[i] if inject else None
            if torch_rng and not probability_flow:
                nz = torch.randn_like(x)  # reference order: drawn BEFORE the net call (sde_helper2.py:47)
            r = None if (nz is not None or probability_flow) else _rng.next(draw_dev)
            score = _call_score(model, x, t_vec, z_cond)
            return _predictor_kernel(sde, x, score, t_vec, noise=nz, rng=r, probability_flow=probability_flow,
                                     impute=impute, want_mean=want_mean)

        def langevin(x, impute, want_mean):
            xm = x
            for k in range(n_steps):
                grad = _call_score(model, x, t_vec, z_cond)
                nz = noise_corr[i, k] if inject else None
                if torch_rng:
                    nz = torch.randn_like(x)  # reference order: drawn AFTER the net call (sde_helper2.py:96)
                r = None if nz is not None else _rng.next(draw_dev)
                final_k = k == n_steps - 1
                x, xm = _corrector_kernels(sde, x, grad, t_vec, target_snr, noise=nz, rng=r,
                                           impute=impute if final_k else None, want_mean=want_mean and final_k,
                                           global_batch=global_batch, reduce_fn=reduce_fn, acc=acc)
            return x, xm

        if predictor_first:
            if pc:
                x, _ = predictor(x, None, False)
                return langevin(x, im, last)
            return predictor(x, im, last)
        if pc:
            x, _ = langevin(x, None, False)
        return predictor(x, im, last)

    if use_graph and steps > 2 and not inject and not torch_rng:
        x, x_mean = _graph_replay(one_step, x, steps, ts, t_vec, n_draws=(1 if not probability_flow else 0) +
                                  (n_steps if pc else 0))
    else:
        for i in range(steps):
            last = i == steps - 1
            x, xm = one_step(i, x, last=last)
            if last:
                x_mean = xm
    out = x_mean
    if conditional:
        ls = _latent_shape(out)
        sc = sde._c()
        imf = _impute_struct(z_obs, obs_mask, False, 0.0)
        L.check(L.lib().sbm_impute_observed(C.byref(ls), C.byref(sc), L.ptr(out), L.ptr(out), C.byref(imf),
                                            L.stream_ptr()), "sbm_impute_observed")
    return (out, x) if return_state else out


def _graph_replay(one_step, x, steps, ts, t_vec, n_draws):
    """Capture ONE predictor-corrector step (2 net forwards + the fused sampler kernels) into a CUDA graph and replay
    it for steps 0..steps-2; every per-step quantity (t, t_next, Philox draw id) lives in device memory and is
    advanced by the sbm_sampler_tick kernel inside the graph.  The last step runs eagerly (it alone writes x_mean
    and skips the imputation)."""
    dev = x.device
    B = x.shape[0]
    step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    draw_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    t_next_dev = torch.zeros(1, dtype=torch.float32, device=dev)
    x_static = x.clone()
    base_draw = _rng.draw

    def tick(advance):
        L.check(L.lib().sbm_sampler_tick(L.ptr(ts), C.c_int32(ts.numel()), L.ptr(step_dev), L.ptr(draw_dev),
                                         L.ptr(t_vec), C.c_int32(B), L.ptr(t_next_dev), C.c_int32(advance),
                                         C.c_uint64(n_draws), L.stream_ptr()), "sbm_sampler_tick")

    def body():
        tick(0)
        _rng.draw = base_draw  # the device counter supplies the per-step offset
        xn, _ = one_step(0, x_static, last=False, graph_state=(t_next_dev, draw_dev))
        x_static.copy_(xn)
        tick(1)

    # warm-up on a side stream (packs weights, sizes the allocator), then restore the state and capture
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    x_static.copy_(x)
    step_dev.zero_()
    draw_dev.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    x_static.copy_(x)
    step_dev.zero_()
    draw_dev.zero_()
    for _ in range(steps - 1):
        g.replay()
    # final step, eager
    _rng.draw = base_draw + (steps - 1) * n_draws
    x_fin, x_mean = one_step(steps - 1, x_static, last=True)
    return x_fin, x_mean


def uncond_sampler(sample_shape, model, device, sde, eps=1e-3, probability_flow=False, pc=False, n_steps=1,
                   target_snr=0.16, cl_g=None, cl_s=None, target=None, *, rng="philox", use_graph=False):
    """sde_helper2.py:115-128: prior -> N x (corrector if pc; predictor) -> x_mean."""
    if cl_g is not None:
        raise NotImplementedError("guidance is not part of the B200 path (see _reject_guidance)")
    device = torch.device(device)
    if rng == "torch":
        x = sde.prior_sampling(sample_shape).to(device)  # CPU draw + H2D, exactly like the reference (:118)
        return pc_sampler(x, model, sde, eps=eps, pc=pc, n_steps=n_steps, target_snr=target_snr,
                          predictor_first=False, probability_flow=probability_flow, rng="torch")
    x = randn(sample_shape, device, scale=float(sde.sigma_max) if isinstance(sde, VESDE) else 1.0)
    return pc_sampler(x, model, sde, eps=eps, pc=pc, n_steps=n_steps, target_snr=target_snr, predictor_first=False,
                      probability_flow=probability_flow, use_graph=use_graph)


def randn(shape, device, scale=1.0):
    """Standard normals from the in-kernel Philox stream (device-side replacement of sde.prior_sampling)."""
    out = torch.empty(tuple(shape), device=device, dtype=torch.float32)
    n = out.numel()
    per_sample = n // shape[0]
    r = _rng.next()
    L.check(L.lib().sbm_randn(L.ptr(out), C.c_int64(n), C.c_uint64(r.seed), C.c_uint64(r.draw),
                              C.c_uint64(r.sample_offset * per_sample), C.c_float(scale), L.stream_ptr()),
            "sbm_randn")
    return out


def cond_sampler(z_obs, given, all_mods, model, sde, eps=1e-3, noise_obs=True, pc=True, n_steps=1, target_snr=0.16,
                 pc_order="predictor_first", probability_flow=False, *, x_init=None, dim=None, use_graph=False,
                 global_batch=None, reduce_fn=None, noise_pred=None, noise_corr=None, num_steps=None, rng="philox",
                 z_cond=None):
    """Conditional generation: sample the missing modalities given the observed ones.

    z_obs   : dict {mod: [B, size_z]} of clean encoder latents for the observed modalities (the reference's
              `z[mod]`), or an already stacked [B, M, D, D] tensor (only the `given` channels are read);
    given   : string of observed modality keys (e.g. '0', '12'); all_mods: string of all keys in channel order
    Returns the stacked latent [B, M, D, D]: missing channels = final x_mean, observed channels = clean latents
    (train_lat_celebhq_unet_cont2.py:314-316)."""
    mask = _obs_mask_from(given, all_mods)
    if isinstance(z_obs, dict):
        some = next(iter(z_obs.values()))
        b, size_z = some.shape[0], some.shape[-1]
        d = dim or int(round(size_z ** 0.5))
        dev = some.device
        stacked = torch.zeros((b, len(all_mods), d, d), device=dev, dtype=torch.float32)
        for i, m in enumerate(all_mods):
            if m in z_obs and m in given:
                stacked[:, i] = z_obs[m].reshape(b, d, d)
    else:
        stacked = _f32c(z_obs)
        dev = stacked.device
    _need_cuda(stacked)
    if x_init is None:
        # prior for the missing modalities (train_lat_celebhq_unet_cont2.py:284), drawn on the device
        x_init = randn(stacked.shape, dev, scale=float(sde.sigma_max) if isinstance(sde, VESDE) else 1.0)
    return pc_sampler(x_init, model, sde, z_obs=stacked, obs_mask=mask, eps=eps, noise_obs=noise_obs, pc=pc,
                      n_steps=n_steps, target_snr=target_snr, predictor_first=(pc_order == "predictor_first"),
                      probability_flow=probability_flow, use_graph=use_graph, global_batch=global_batch,
                      reduce_fn=reduce_fn, noise_pred=noise_pred, noise_corr=noise_corr, num_steps=num_steps, rng=rng,
                      z_cond=z_cond)


# ======================================================================================= DSM loss
class _DsmLossFn(torch.autograd.Function):
    """loss = mean_b red_CHW(term^2)[*g2]; backward = grad_out * d loss / d score (computed by the same kernel)."""

    @staticmethod
    def forward(ctx, score, z, std, g2, lw, reduce_mean, global_batch):
        ls = _latent_shape(score)
        score = score.contiguous()
        dscore = torch.empty_like(score)
        acc = torch.zeros(1, dtype=torch.float64, device=score.device)
        L.check(L.lib().sbm_dsm_loss(C.byref(ls), L.ptr(score), L.ptr(z), L.ptr(std), L.ptr(g2), L.ptr(dscore),
                                     L.ptr(acc), C.c_int32(1 if lw else 0), C.c_int32(1 if reduce_mean else 0),
                                     C.c_int64(global_batch), L.stream_ptr()), "sbm_dsm_loss")
        loss = torch.empty((), dtype=torch.float32, device=score.device)
        L.check(L.lib().sbm_f64_to_f32(L.ptr(acc), L.ptr(loss), C.c_int32(1), L.stream_ptr()), "sbm_f64_to_f32")
        ctx.save_for_backward(dscore)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dscore,) = ctx.saved_tensors
        g = torch.empty_like(dscore)
        go = grad_out.contiguous().float()
        L.check(L.lib().sbm_scale_by_scalar(L.ptr(dscore), L.ptr(go), L.ptr(g), C.c_int64(dscore.numel()),
                                            L.stream_ptr()), "sbm_scale_by_scalar")
        return g, None, None, None, None, None, None


def loss_fn(batch, score_fn, sde, reduce_mean=True, likelihood_weighting=True, eps=1e-5, im_sample=False, *,
            u=None, z=None, rng="torch", global_batch=None, draw_dev=None, z_cond=None):
    """Denoising-score-matching loss (sde_helper2.py:152-186) -> 0-d tensor, differentiable w.r.t. the score net.
    Two fused kernels around the net call: perturb (t, x~ = mean + std z) and loss (+ its gradient)."""
    _need_cuda(batch)
    batch = _f32c(batch)
    B = batch.shape[0]
    dev = batch.device
    if likelihood_weighting and im_sample:
        # importance-sampled t (sde_helper2.py:131-150, 164-165): host-side numpy bisection, as in the reference
        u_t = torch.tensor(np.array(sample_importance_weighted_time_for_likelihood(B, sde.beta_0, sde.beta_1,
                                                                                   T=sde.T)), dtype=torch.float32)
        u = ((u_t - eps) / (sde.T - eps)).to(dev)
    r = None
    if u is None or z is None:
        if rng == "torch":
            u = torch.rand(B, device=dev) if u is None else u      # t first (:167) ...
            z = torch.randn_like(batch) if z is None else z        # ... then z (:168)
        else:
            r = _rng.next(draw_dev)  # draw_dev: device-side draw-id offset (CUDA-graph replay of a training step)
            _rng.draw += 1  # perturb consumes two draw ids (u, z)
    ls = _latent_shape(batch)
    sc = sde._c()
    xt = torch.empty_like(batch)
    z_out = torch.empty_like(batch)
    t = torch.empty(B, device=dev, dtype=torch.float32)
    std = torch.empty(B, device=dev, dtype=torch.float32)
    lw_branch = bool(likelihood_weighting and not im_sample)
    g2 = torch.empty(B, device=dev, dtype=torch.float32) if lw_branch else None
    L.check(L.lib().sbm_dsm_perturb(C.byref(ls), C.byref(sc), L.ptr(batch), L.ptr(u), L.ptr(z), L.ptr(xt),
                                    L.ptr(z_out), L.ptr(t), L.ptr(std), L.ptr(g2), C.c_float(eps),
                                    C.byref(r) if r is not None else None, L.stream_ptr()), "sbm_dsm_perturb")
    score = _call_score(score_fn, xt, t, z_cond)
    return _DsmLossFn.apply(score.float(), z_out, std, g2, lw_branch, bool(reduce_mean), global_batch or B)


# importance-sampled time for the likelihood-weighted loss: host-side numpy, not accelerated (off in every shipped
# command: --ll-weighting=0).  sde_helper2.py:131-150 with jax.numpy -> numpy.
def likelihood_importance_cum_weight(t, beta_0, beta_1, eps=1e-5):
    e1 = 0.5 * eps * (eps - 2) * beta_0 - 0.5 * eps ** 2 * beta_1
    e2 = 0.5 * t * (t - 2) * beta_0 - 0.5 * t ** 2 * beta_1
    term1 = np.where(np.abs(e1) <= 1e-3, -e1, 1.0 - np.exp(e1))
    term2 = np.where(np.abs(e2) <= 1e-3, -e2, 1.0 - np.exp(e2))
    return 0.5 * (-2 * np.log(term1) + 2 * np.log(term2) + beta_0 * (-2 * eps + eps ** 2 - (t - 2) * t)
                  + beta_1 * (-eps ** 2 + t ** 2))


def sample_importance_weighted_time_for_likelihood(shape, beta_0, beta_1, quantile=None, eps=1e-5, steps=100, T=1):
    Z = likelihood_importance_cum_weight(T, beta_0, beta_1, eps)
    if quantile is None:
        quantile = torch.distributions.uniform.Uniform(0, float(Z)).sample((shape,)).numpy()
    lb = np.ones_like(quantile) * eps
    ub = np.ones_like(quantile) * T
    for _ in range(steps):
        mid = (lb + ub) / 2.0
        value = likelihood_importance_cum_weight(mid, beta_0, beta_1, eps=eps)
        lb = np.where(value <= quantile, mid, lb)
        ub = np.where(value <= quantile, ub, mid)
    return (lb + ub) / 2.0
