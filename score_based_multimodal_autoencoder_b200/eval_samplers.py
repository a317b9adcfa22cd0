"""Evaluation plumbing either side of the hot path.

* `annealed_langevin_sampler`: the legacy NCSN-style evaluator of `eval_lat_celeba_hq_all.py:244-279` (500 noise levels
  sigma = linspace(5, 0.1, 500), `n_comp` Langevin steps per level, per-modality step sizes `er[mod]` and noise scales
  `c[mod]`, a score net conditioned on the integer noise-level index);
* `langevin_refine`: the fixed-step evaluator of `fid_upd10.py:271-290` (`z += lr1 [* (i+1)/n_comp] * s(z) + lr2 * randn`,
  a score net without time input over the flattened latents);
* `save_checkpoint` / `load_checkpoint`: the reference's checkpoint container
  `{'epoch', 'model_state_dict', 'train_loss', 'val_loss', 'size_z'}` (train_lat_celebhq_unet_cont2.py:534-557, :480).

The per-step update of ALL missing modalities is one fused kernel (`sbm_langevin_axpy_step`, 12 B / latent element; the
reference runs 5 small kernels per missing modality per step and re-stacks the latents each time); the score net is the
caller's (`Unet` / `UNetModel` of this package, or any callable returning a CUDA tensor).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import sde_helper2 as sh


def _stack(z, all_mods, dim=None):
    if isinstance(z, dict):
        some = next(iter(z.values()))
        b, size_z = some.shape[0], some.shape[-1]
        d = dim or int(round(size_z ** 0.5))
        return torch.cat([z[m].reshape(b, 1, d, d) for m in all_mods], dim=1).float().contiguous()
    return z.detach().float().contiguous()


def _axpy_step(x, score, coef_a, coef_b, obs_mask, *, noise=None, rng=None, out=None):
    ls = sh._latent_shape(x)
    m = ls.mods
    out = out if out is not None else torch.empty_like(x)
    a = (C.c_float * m)(*[float(v) for v in coef_a])
    b = (C.c_float * m)(*[float(v) for v in coef_b])
    L.check(L.lib().sbm_langevin_axpy_step(C.byref(ls), L.ptr(x), L.ptr(score), L.ptr(noise), a, b,
                                           C.c_uint32(obs_mask), L.ptr(out), C.byref(rng) if rng is not None else None,
                                           L.stream_ptr()), "sbm_langevin_axpy_step")
    return out


@torch.no_grad()
def annealed_langevin_sampler(z, given, all_mods, model, er, c, n_comp=1, sigmas=None, *, noise=None, rng="philox",
                              num_levels=None):
    """eval_lat_celeba_hq_all.py:258-275.

    z        : dict {mod: [B, size_z]} (observed modalities = encoder latents, missing = N(0,1) prior draws, :247-256)
               or the stacked [B, M, D, D] tensor;
    er, c    : dicts {mod: float} of step sizes and noise scales (:471-511);
    sigmas   : noise levels, default np.linspace(5, 0.1, 500) (:222); the net is called as model(z_all, level_index)
               with the integer level index as its conditioning input and its output is divided by sigma (:270);
    noise    : optional [levels, n_comp, B, M, D, D] injected normals (parity tests), else torch / in-kernel Philox.
    Returns the stacked latent [B, M, D, D]; observed channels are returned untouched."""
    x = _stack(z, all_mods)
    sh._need_cuda(x)
    b = x.shape[0]
    sig = np.linspace(5, 0.1, 500) if sigmas is None else np.asarray(sigmas, dtype=np.float64)
    mask = sh._obs_mask_from(given, all_mods)
    levels = len(sig) if num_levels is None else num_levels
    for s_in in range(levels):
        idx = torch.full((b,), s_in, device=x.device, dtype=torch.long)
        sigma = float(np.float32(sig[s_in]))                  # `cur_sigmas = sigmas[sigma_index].float()` (:260)
        ratio = sig[s_in] ** 2 / sig[-1] ** 2                 # float64, like the reference's 0-dim tensor (:274)
        alpha = [er[m] * ratio for m in all_mods]
        coef_a = [a / sigma for a in alpha]                   # alpha * (model_out / sigma)
        coef_b = [c[m] * float(np.sqrt(2 * a)) for m, a in zip(all_mods, alpha)]
        for i in range(n_comp):
            score = model(x, idx).float().contiguous()
            nz, r = None, None
            if noise is not None:
                nz = noise[s_in, i]
            elif rng == "torch":
                nz = torch.randn_like(x)
            else:
                r = sh._rng.next()
            x = _axpy_step(x, score, coef_a, coef_b, mask, noise=nz, rng=r)
    return x


@torch.no_grad()
def langevin_refine(z, predicted_mods, all_mods, sm_model, n_comp, lr1, lr2, schedule=False, *, noise=None,
                    rng="philox", dim=None):
    """fid_upd10.py:279-290: `n_comp` fixed-step Langevin updates of the predicted modalities with a time-free score
    net over the FLATTENED latents, `sm_out = sm_model(cat(z[mod] for mod in all_mods))` -> [B, M * size_z].
    schedule=True scales the step by (i+1)/n_comp and adds one more unit-step update at the end (:287-288).
    noise: optional [n_comp (+1 with schedule), B, M, D, D]."""
    x = _stack(z, all_mods, dim)
    sh._need_cuda(x)
    b, m = x.shape[0], x.shape[1]
    given = "".join(k for k in all_mods if k not in predicted_mods)
    mask = sh._obs_mask_from(given, all_mods)
    draws = 0

    def update(x, score, a):
        nonlocal draws
        nz, r = None, None
        if noise is not None:
            nz = noise[draws]
        elif rng == "torch":
            nz = torch.randn_like(x)
        else:
            r = sh._rng.next()
        draws += 1
        return _axpy_step(x, score, [a] * m, [lr2] * m, mask, noise=nz, rng=r)

    for i in range(n_comp):
        score = sm_model(x.view(b, -1)).float().contiguous().view_as(x)
        if not schedule:
            x = update(x, score, lr1)
        else:
            x = update(x, score, lr1 * ((i + 1) / n_comp))
            if i == n_comp - 1:
                x = update(x, score, 1 * ((i + 1) / n_comp))      # the reference re-uses sm_out of this iteration
    return x


# ------------------------------------------------------------------------------------------ checkpoint container
def save_checkpoint(path, model, epoch, train_loss=None, val_loss=None, size_z=None, **extra):
    """train_lat_celebhq_unet_cont2.py:540-546 / 552-558: the dict the reference's trainers write (and :480 reads).
    `model.state_dict()` of this package's nets has the reference's key names and shapes, so the
    file loads into the reference's modules and vice versa."""
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.save({"epoch": epoch, "model_state_dict": sd, "train_loss": train_loss, "val_loss": val_loss,
                "size_z": size_z, **extra}, path)


def load_checkpoint(path, model, map_location="cpu", strict=True):
    """`model.load_state_dict(torch.load(path, map_location=device)['model_state_dict'])` (:480); returns the rest of
    the container (epoch, losses, size_z)."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ck["model_state_dict"], strict=strict)
    return {k: v for k, v in ck.items() if k != "model_state_dict"}
