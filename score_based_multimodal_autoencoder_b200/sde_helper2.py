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

    def _disc_table_on(self, device):
        """Device copy of the table `discretize` indexes: discrete_betas (VPSDE) / discrete_sigmas (VESDE); None for
        subVPSDE, which has no override and takes the base class's Euler-Maruyama rule (sde_helper2.py:236-253)."""
        src = getattr(self, "discrete_sigmas", None) if self._kind == "ve" else (
            self.discrete_betas if self._kind == "vp" else None)
        if src is None:
            return None
        key = ("disc", str(device))
        tab = self._dev_tables.get(key)
        if tab is None:
            tab = self._dev_tables[key] = src.to(device=device, dtype=torch.float32).contiguous()
        return tab

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
        drift, diffusion = self._fwd.sde(x, t)
        score = _guided(self._score_fn(x, t), x, t, cl_g, cl_s, given, all_mods)
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


_PREDICTORS = ("euler", "reverse_diffusion")


def _predictor_kernel(sde, x, score, t, *, noise=None, rng=None, probability_flow=False, impute=None, want_mean=True,
                      out=None, predictor="euler"):
    """One fused launch: Euler-Maruyama (`sbm_predictor_step`) or reverse-diffusion (`sbm_rd_predictor_step`) rule."""
    ls = _latent_shape(x)
    x_new = out if out is not None else torch.empty_like(x)
    x_mean = torch.empty_like(x) if want_mean else None
    sc = sde._c()
    common = (L.ptr(noise), L.ptr(x_new), L.ptr(x_mean), C.c_int32(1 if probability_flow else 0),
              C.byref(rng) if rng is not None else None, C.byref(impute) if impute is not None else None,
              L.stream_ptr())
    if predictor == "euler":
        L.check(L.lib().sbm_predictor_step(C.byref(ls), C.byref(sc), L.ptr(x), L.ptr(score), L.ptr(t), *common),
                "sbm_predictor_step")
    elif predictor == "reverse_diffusion":
        L.check(L.lib().sbm_rd_predictor_step(C.byref(ls), C.byref(sc), L.ptr(x), L.ptr(score), L.ptr(t),
                                              L.ptr(sde._disc_table_on(x.device)), *common), "sbm_rd_predictor_step")
    else:
        raise ValueError(f"predictor must be one of {_PREDICTORS}, got {predictor!r}")
    return x_new, x_mean


_side_streams: dict = {}


def _side_stream(dev):
    s = _side_streams.get(dev.index)
    if s is None:
        s = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _fork_noise_norm(x, rng, acc):
    """acc[1] += sum_b ||noise_b|| of the Philox draw `rng`, on a side stream: the draw depends on no data, so it runs
    beside the score-net forward that produces the other operand of the corrector's step size (sde_helper2.py:96-99).
    Returns the stream to join (`torch.cuda.current_stream().wait_stream(side)`) before the norms are consumed."""
    ls = _latent_shape(x)
    side = _side_stream(x.device)
    side.wait_stream(torch.cuda.current_stream())
    L.check(L.lib().sbm_noise_norm(C.byref(ls), C.byref(rng), L.ptr(acc), C.c_void_p(side.cuda_stream)),
            "sbm_noise_norm")
    return side


def _corrector_kernels(sde, x, grad, t, target_snr, *, noise=None, rng=None, impute=None, want_mean=True,
                       global_batch=None, reduce_fn=None, acc=None, out=None, noise_norm_done=False, join=None):
    """norms kernel -> (optional cross-rank sum of the two batch norms) -> update kernel.  `acc` is a zeroed buffer of
    3 doubles (2 sums + a completion ticket); the update kernel re-zeroes it, so a reused `acc` never needs a memset.
    `noise_norm_done`: acc[1] gets the Philox noise norm from `_fork_noise_norm`, the norms kernel reads the score only;
    `join` = that side stream: it is joined AFTER the score-norm kernel has been issued (the two kernels add into
    different slots of `acc`), so the noise kernel can also run beside it -- only the update kernel needs both sums."""
    ls = _latent_shape(x)
    if acc is None:
        acc = torch.zeros(3, dtype=torch.float64, device=x.device)
    elif acc.numel() < 3:
        raise L.SbmError("corrector accumulator needs 3 doubles (2 sums + ticket)")
    rp = C.byref(rng) if rng is not None else None
    L.check(L.lib().sbm_corrector_norms(C.byref(ls), L.ptr(grad), L.ptr(noise), None if noise_norm_done else rp,
                                        L.ptr(acc), L.stream_ptr()), "sbm_corrector_norms")
    if join is not None:
        torch.cuda.current_stream().wait_stream(join)
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


def _guided(score, x, t, cl_g, cl_s, given, all_mods):
    """Classifier / EBM guidance of the score (sde_helper2.py:65-94, 283-312); a no-op without `cl_g` or `given`."""
    if cl_g is None or given is None or not given:
        return score
    from .guidance import apply_guidance
    return apply_guidance(score, x, t, cl_g, cl_s, given, all_mods)


def em_predictor(x, t, score_fn, sde, probability_flow=False, cl_g=None, cl_s=None, target=None, given=None,
                 all_mods=None, z_cond=None, *, noise=None, rng="torch"):
    """Euler-Maruyama reverse-SDE predictor step (sde_helper2.py:45-52) -> (x, x_mean).
    One fused kernel after the score-net call; the noise is drawn BEFORE the net call like the reference."""
    _need_cuda(x, t)
    x, t = _f32c(x), _f32c(t)
    r = None
    if noise is None and not probability_flow:
        if rng == "torch":
            noise = torch.randn_like(x)
        else:
            r = _rng.next()
    score = _f32c(_call_score(score_fn, x, t, z_cond))
    score = _guided(score, x, t, cl_g, cl_s, given, all_mods)
    return _predictor_kernel(sde, x, score, t, noise=noise, rng=r, probability_flow=probability_flow)


def rd_predictor(x, t, score_fn, sde, probability_flow=False, z_cond=None, *, noise=None, rng="torch"):
    """Reverse-diffusion (ancestral) predictor step -> (x, x_mean):
        (rev_f, rev_G) = sde.reverse(score_fn, probability_flow).discretize(x, t);  x_mean = x - rev_f;
        x = x_mean + rev_G * z.
    The reference ships this discretisation (sde_helper2.py:236-253, 319-324, 373-381, 465-473) without a caller
    (SURVEY.md 8a-6); here it is a mode of the fused predictor kernel (`sbm_rd_predictor_step`, same 12 B / element)
    and selectable in the N-step samplers with `predictor="reverse_diffusion"`."""
    _need_cuda(x, t)
    x, t = _f32c(x), _f32c(t)
    r = None
    if noise is None and not probability_flow:
        if rng == "torch":
            noise = torch.randn_like(x)     # drawn before the net call, like em_predictor (:47)
        else:
            r = _rng.next()
    score = _f32c(_call_score(score_fn, x, t, z_cond))
    return _predictor_kernel(sde, x, score, t, noise=noise, rng=r, probability_flow=probability_flow,
                             predictor="reverse_diffusion")


def corrector(x, t, score_fn, sde, n_steps, target_snr, cl_g=None, cl_s=None, target=None, given=None, all_mods=None,
              z_cond=None, *, noise=None, rng="torch", global_batch=None, reduce_fn=None):
    """Langevin corrector (sde_helper2.py:54-106) -> (x, x_mean).  Two fused kernels per Langevin step
    (batch-coupled norms, then the update); the noise is drawn AFTER the net call like the reference.
    `noise`: optional [n_steps, B, M, D, D] injected noise."""
    _need_cuda(x, t)
    x, t = _f32c(x), _f32c(t)
    x_mean = x
    acc = torch.zeros(3, dtype=torch.float64, device=x.device)
    for i in range(n_steps):
        nz, r, side = None, None, None
        if noise is not None:
            nz = noise[i] if noise.dim() == 5 else noise
        elif rng != "torch":
            r = _rng.next()
            side = _fork_noise_norm(x, r, acc)
        grad = _f32c(_call_score(score_fn, x, t, z_cond))
        grad = _guided(grad, x, t, cl_g, cl_s, given, all_mods)
        if nz is None and r is None:
            nz = torch.randn_like(x)
        x, x_mean = _corrector_kernels(sde, x, grad, t, target_snr, noise=nz, rng=r, global_batch=global_batch,
                                       reduce_fn=reduce_fn, acc=acc, noise_norm_done=side is not None, join=side)
    return x, x_mean


# ======================================================================================= N-step samplers
def _obs_mask_from(given, all_mods) -> int:
    mask = 0
    if given:
        for i, m in enumerate(all_mods):
            if m in given:
                mask |= 1 << i
    return mask


class _PCRun:
    """The per-step arithmetic of one predictor-corrector run over fixed buffers.

    Semantics = the reference's inline loop (train_lat_celebhq_unet_cont2.py:287-316 for predictor_first=True,
    :173-200 / sde_helper2.py:121-128 for predictor_first=False).  Per step: 2 score-net forwards + the fused
    elementwise kernels (predictor, corrector norms, corrector update; the Philox noise norm beside the net on a side
    stream); the imputation of the observed channels for step i+1 is an epilogue of the last kernel of step i.
    `z_obs` / `z_cond` are the buffers the steps READ (static copies when the step is captured into a CUDA graph)."""

    def __init__(self, model, sde, B, dev, *, z_obs, obs_mask, z_cond, eps, noise_obs, pc, n_steps, target_snr,
                 predictor_first, probability_flow, predictor, global_batch, reduce_fn, guidance=None):
        if predictor not in _PREDICTORS:
            raise ValueError(f"predictor must be one of {_PREDICTORS}, got {predictor!r}")
        self.model, self.sde, self.dev, self.B = model, sde, dev, B
        self.z_obs, self.obs_mask, self.z_cond = z_obs, obs_mask, z_cond
        self.conditional = z_obs is not None and obs_mask != 0
        self.noise_obs, self.pc, self.n_steps, self.target_snr = noise_obs, pc, n_steps, target_snr
        self.predictor_first, self.probability_flow, self.predictor = predictor_first, probability_flow, predictor
        self.global_batch, self.reduce_fn, self.guidance = global_batch, reduce_fn, guidance
        self.N = sde.N
        self.ts = torch.linspace(sde.T, eps, sde.N, device=dev)
        self.ts_host = self.ts.tolist()
        self.t_vec = torch.empty(B, device=dev, dtype=torch.float32)
        self.acc = torch.zeros(3, dtype=torch.float64, device=dev)
        self.n_draws = (0 if probability_flow else 1) + (n_steps if pc else 0)

    def _score(self, x):
        with L.nvtx("sbm.score_net"):
            s = _call_score(self.model, x, self.t_vec, self.z_cond)
        if self.guidance is not None:
            cl_g, cl_s, given, all_mods = self.guidance
            s = _guided(_f32c(s), x, self.t_vec, cl_g, cl_s, given, all_mods)
        return s

    def impute(self, x, t, noise_obs):
        """x[:, observed] <- (noised) clean latents, in place (first step / finishing)."""
        if not self.conditional:
            return
        ls, sc = _latent_shape(x), self.sde._c()
        im = _impute_struct(self.z_obs, self.obs_mask, noise_obs, t)
        L.check(L.lib().sbm_impute_observed(C.byref(ls), C.byref(sc), L.ptr(x), L.ptr(x), C.byref(im),
                                            L.stream_ptr()), "sbm_impute_observed")

    def step(self, i, x, *, last, graph_state=None, noise_pred=None, noise_corr=None, torch_rng=False):
        """Runs step i on x -> (x, x_mean or None).  `graph_state` = (t_next_dev, draw_dev) when the step is being
        captured for replay (t_vec is then maintained by sbm_sampler_tick)."""
        t_next_dev, draw_dev = graph_state if graph_state is not None else (None, None)
        t_vec, sde, pf = self.t_vec, self.sde, self.probability_flow
        if graph_state is None:
            t_vec.fill_(self.ts_host[i])
        im = None
        if self.conditional and not last:
            im = _impute_struct(self.z_obs, self.obs_mask, self.noise_obs, self.ts_host[min(i + 1, self.N - 1)],
                                t_next_dev)
        inject = noise_pred is not None

        def predictor(x, impute, want_mean):
            nz = noise_pred[i] if inject else None
            if torch_rng and not pf:
                nz = torch.randn_like(x)  # reference order: drawn BEFORE the net call (sde_helper2.py:47)
            r = None if (nz is not None or pf) else _rng.next(draw_dev)
            score = self._score(x)
            return _predictor_kernel(sde, x, score, t_vec, noise=nz, rng=r, probability_flow=pf, impute=impute,
                                     want_mean=want_mean, predictor=self.predictor)

        def langevin(x, impute, want_mean):
            xm = x
            for k in range(self.n_steps):
                nz = noise_corr[i, k] if inject else None
                r, side = None, None
                if nz is None and not torch_rng:
                    r = _rng.next(draw_dev)
                    side = _fork_noise_norm(x, r, self.acc)
                grad = self._score(x)
                if torch_rng:
                    nz = torch.randn_like(x)  # reference order: drawn AFTER the net call (sde_helper2.py:96)
                final_k = k == self.n_steps - 1
                x, xm = _corrector_kernels(sde, x, grad, t_vec, self.target_snr, noise=nz, rng=r,
                                           impute=impute if final_k else None, want_mean=want_mean and final_k,
                                           global_batch=self.global_batch, reduce_fn=self.reduce_fn, acc=self.acc,
                                           noise_norm_done=side is not None, join=side)
            return x, xm

        with L.nvtx(f"sbm.pc_step[{i}]"):
            if self.predictor_first:
                if self.pc:
                    x, _ = predictor(x, None, False)
                    return langevin(x, im, last)
                return predictor(x, im, last)
            if self.pc:
                x, _ = langevin(x, None, False)
            return predictor(x, im, last)


@torch.no_grad()
def pc_sampler(x0, model, sde, *, z_obs=None, obs_mask=0, eps=1e-3, noise_obs=True, pc=True, n_steps=1,
               target_snr=0.16, predictor_first=True, probability_flow=False, noise_pred=None, noise_corr=None,
               num_steps=None, global_batch=None, reduce_fn=None, use_graph=False, return_state=False,
               rng="philox", z_cond=None, predictor="euler", cl_g=None, cl_s=None, given=None, all_mods=None):
    """N-step predictor-corrector sampler over a stacked latent [B,M,D,D] with observed-modality imputation
    (see `_PCRun`).  Returns the last x_mean with the observed channels set to the clean latents.

    predictor : "euler" (em_predictor, the reference's sampler) or "reverse_diffusion" (the discretize() rule);
    use_graph : replay ONE captured CUDA graph per step; the capture is cached across calls (same model weights,
                shapes and options), so repeated calls cost a copy-in plus the replays;
    cl_g, cl_s, given, all_mods : classifier / EBM guidance of both score evaluations (sde_helper2.py:65-94, 283-312)."""
    _need_cuda(x0)
    dev = x0.device
    B = x0.shape[0]
    steps = sde.N if num_steps is None else num_steps
    inject = noise_pred is not None
    torch_rng = (rng == "torch") and not inject
    guidance = (cl_g, cl_s, given, all_mods) if (cl_g is not None and given) else None
    opts = dict(obs_mask=obs_mask if z_obs is not None else 0, eps=eps, noise_obs=noise_obs, pc=pc, n_steps=n_steps,
                target_snr=target_snr, predictor_first=predictor_first, probability_flow=probability_flow,
                predictor=predictor, global_batch=global_batch, reduce_fn=reduce_fn, guidance=guidance)
    if use_graph and steps >= 1 and not inject and not torch_rng and guidance is None:
        return _graph_cache.run(model, sde, x0, z_obs, z_cond, steps, opts, return_state)
    x = _f32c(x0).clone()
    run = _PCRun(model, sde, B, dev, z_obs=_f32c(z_obs) if opts["obs_mask"] else None,
                 z_cond=z_cond, **opts)
    run.impute(x, run.ts_host[0], noise_obs)
    x_mean = x
    for i in range(steps):
        last = i == steps - 1
        x, xm = run.step(i, x, last=last, noise_pred=noise_pred, noise_corr=noise_corr, torch_rng=torch_rng)
        if last:
            x_mean = xm
    out = x_mean
    run.impute(out, 0.0, False)
    return (out, x) if return_state else out


class _PCGraph:
    """ONE predictor-corrector step (2 net forwards + the fused sampler kernels) captured as a CUDA graph over static
    buffers, plus the variant for the final step (writes x_mean, skips the imputation, finishes with the clean
    observed latents).  Every per-step quantity (t, t_next, Philox draw id) lives in device memory and is advanced by
    the sbm_sampler_tick kernel inside the graph, so the same graph serves all N steps (graph replay equals the eager
    loop bit for bit -- tests/test_sampler_gpu.py)."""

    def __init__(self, model, sde, x0, z_obs, z_cond, opts):
        dev = x0.device
        B = x0.shape[0]
        self.x = torch.empty_like(x0, dtype=torch.float32).contiguous()
        self.out = torch.empty_like(self.x)
        self.z_obs = torch.empty_like(self.x) if opts["obs_mask"] else None
        self.z_cond = torch.empty_like(z_cond) if z_cond is not None else None
        self.run = _PCRun(model, sde, B, dev, z_obs=self.z_obs, z_cond=self.z_cond, **opts)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.draw_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.t_next_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.graphs = {}
        self.launches = {}
        self.seed = (_rng.seed, _rng.sample_offset)

    def _tick(self, advance):
        r = self.run
        L.check(L.lib().sbm_sampler_tick(L.ptr(r.ts), C.c_int32(r.ts.numel()), L.ptr(self.step_dev),
                                         L.ptr(self.draw_dev), L.ptr(r.t_vec), C.c_int32(r.B), L.ptr(self.t_next_dev),
                                         C.c_int32(advance), C.c_uint64(r.n_draws), L.stream_ptr()),
                "sbm_sampler_tick")

    def _body(self, last):
        self._tick(0)
        host_draw = _rng.draw
        _rng.draw = 0  # the device counter carries the whole draw id
        try:
            xn, xm = self.run.step(0, self.x, last=last, graph_state=(self.t_next_dev, self.draw_dev))
        finally:
            _rng.draw = host_draw
        self.x.copy_(xn)
        if last:
            self.out.copy_(xm)
            self.run.impute(self.out, 0.0, False)
        self._tick(1)

    def _capture(self, last):
        dev = self.x.device
        keep = (self.x.clone(), self.step_dev.clone(), self.draw_dev.clone())
        # warm-up on a side stream (packs weights, sizes the allocator), then restore the state and capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body(last)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize(dev)
        n0 = L.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._body(last)
        self.launches[last] = L.launch_count() - n0
        self.x.copy_(keep[0])
        self.step_dev.copy_(keep[1])
        self.draw_dev.copy_(keep[2])
        self.graphs[last] = g
        return g

    def sample(self, x0, z_obs, z_cond, steps, return_state):
        r = self.run
        self.x.copy_(x0)
        if self.z_obs is not None:
            self.z_obs.copy_(z_obs)
        if self.z_cond is not None:
            self.z_cond.copy_(z_cond)
        self.step_dev.zero_()
        self.draw_dev.fill_(_rng.draw)
        r.impute(self.x, r.ts_host[0], r.noise_obs)
        if steps > 1:
            g = self.graphs.get(False) or self._capture(False)
            for _ in range(steps - 1):
                g.replay()
        g = self.graphs.get(True) or self._capture(True)
        g.replay()
        _rng.draw += steps * r.n_draws
        out = self.out.clone()
        return (out, self.x.clone()) if return_state else out


class _GraphCache:
    """Captured sampler steps, kept across calls.  An entry is valid for one (model, weight versions, SDE, latent shape,
    sampler options, Philox seed / shard offset); a changed weight re-captures in place.  Bounded (LRU): a captured
    CelebA-sized step owns a private pool of a few GB of activations."""

    max_entries = 4

    def __init__(self):
        self.entries: dict = {}

    @staticmethod
    def _weights_sig(model):
        params = getattr(model, "parameters", None)
        if params is None:
            return None
        return tuple((p.data_ptr(), p._version) for p in params())

    def run(self, model, sde, x0, z_obs, z_cond, steps, opts, return_state):
        import weakref
        x0 = _f32c(x0)
        conditional = bool(opts["obs_mask"])
        key = (id(model), type(sde).__name__, float(sde.beta_0), float(sde.beta_1), int(sde.N), tuple(x0.shape),
               str(x0.device), None if z_cond is None else (tuple(z_cond.shape), z_cond.dtype),
               tuple((k, id(v) if callable(v) else v) for k, v in sorted(opts.items()) if k != "guidance"),
               _rng.seed, _rng.sample_offset)
        sig = self._weights_sig(model)
        hit = self.entries.pop(key, None)
        if hit is not None and (hit[0]() is not model or hit[1] != sig):
            hit = None  # another object at the same address, or the weights changed: capture again
        if hit is None:
            try:
                ref = weakref.ref(model)
            except TypeError:  # plain functions / lambdas used as score nets in tests
                ref = (lambda m: (lambda: m))(model)
            hit = (ref, sig, _PCGraph(model, sde, x0, z_obs if conditional else None, z_cond, opts))
        self.entries[key] = hit  # most recent last
        while len(self.entries) > self.max_entries:
            self.entries.pop(next(iter(self.entries)))
        return hit[2].sample(x0, _f32c(z_obs) if conditional else None, z_cond, steps, return_state)

    def clear(self):
        self.entries.clear()

    def launches_per_step(self):
        """kernels per replayed step of the most recently used entry: {False: mid step, True: final step}"""
        if not self.entries:
            return {}
        return dict(next(reversed(self.entries.values()))[2].launches)


_graph_cache = _GraphCache()


def clear_graph_cache():
    """Drop every cached sampler graph (and the activation pools they own)."""
    _graph_cache.clear()


def uncond_sampler(sample_shape, model, device, sde, eps=1e-3, probability_flow=False, pc=False, n_steps=1,
                   target_snr=0.16, cl_g=None, cl_s=None, target=None, *, rng="philox", use_graph=False,
                   predictor="euler"):
    """sde_helper2.py:115-128: prior -> N x (corrector if pc; predictor) -> x_mean.  `cl_g` is accepted and inert, as
    in the reference: uncond_sampler passes no `given` to its step functions (:125-126), so guidance never fires."""
    device = torch.device(device)
    if rng == "torch":
        x = sde.prior_sampling(sample_shape).to(device)  # CPU draw + H2D, exactly like the reference (:118)
        return pc_sampler(x, model, sde, eps=eps, pc=pc, n_steps=n_steps, target_snr=target_snr,
                          predictor_first=False, probability_flow=probability_flow, rng="torch", predictor=predictor)
    x = randn(sample_shape, device, scale=float(sde.sigma_max) if isinstance(sde, VESDE) else 1.0)
    return pc_sampler(x, model, sde, eps=eps, pc=pc, n_steps=n_steps, target_snr=target_snr, predictor_first=False,
                      probability_flow=probability_flow, use_graph=use_graph, predictor=predictor)


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
                 z_cond=None, predictor="euler", cl_g=None, cl_s=None):
    """Conditional generation: sample the missing modalities given the observed ones.

    z_obs   : dict {mod: [B, size_z]} of clean encoder latents for the observed modalities (the reference's
              `z[mod]`), or an already stacked [B, M, D, D] tensor (only the `given` channels are read);
    given   : string of observed modality keys (e.g. '0', '12'); all_mods: string of all keys in channel order;
    cl_g, cl_s : optional classifier / EBM guidance (dict of pair energies or one index-conditioned energy net and its
              scale, train_lat_celebhq_unet_cont2.py:305-312)
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
                      z_cond=z_cond, predictor=predictor, cl_g=cl_g, cl_s=cl_s, given=given, all_mods=all_mods)


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
