"""Generate tests/golden/dsm_loss_is.pt: the importance-sampled-time branch of the UNMODIFIED reference's loss_fn
(/root/reference/sde_helper2.py:129-150, 164-165, 177-179: likelihood_weighting=True, im_sample=True) with a small
differentiable score function, and check oracle/sde_oracle.py's restatement against it.

Run in the build container only:  python -m oracle.gen_golden_is
"""
from __future__ import annotations

import os

import torch

from . import sde_oracle as so
from .gen_golden import OUT, NoiseFeed, import_reference


def toy_score(w):
    return lambda xx, tt: torch.einsum("oc,bchw->bohw", w, xx) * (1.0 + tt[:, None, None, None].to(xx.dtype))


def main():
    sh, _, _ = import_reference()
    g = torch.Generator().manual_seed(77)
    B = 12
    batch = torch.randn(B, 5, 8, 8, generator=g)
    u01 = torch.rand(B, generator=g)
    z = torch.randn(B, 5, 8, 8, generator=g)
    w0 = torch.randn(5, 5, generator=g) * 0.3
    cases = []
    for a, b, rm in [(0.1, 20.0, True), (1.0, 5.0, False)]:
        sde = sh.VPSDE(a, b, 100)
        spec = so.SdeSpec("vp", a, b, 100)
        w = w0.clone().requires_grad_(True)
        with NoiseFeed([z], [u01]).patched():     # Uniform.sample draws through torch.rand; z through randn_like
            loss = sh.loss_fn(batch, toy_score(w), sde, reduce_mean=rm, likelihood_weighting=True, im_sample=True)
        loss.backward()
        t_or = so.importance_sampled_t(spec, u01)
        with NoiseFeed([], [u01]).patched():
            t_ref = torch.tensor(sh.sample_importance_weighted_time_for_likelihood(B, a, b, T=1))
        assert torch.equal(t_or, t_ref), (t_or, t_ref)
        wo = w0.clone().requires_grad_(True)
        l_or = so.dsm_loss(spec, batch, toy_score(wo), None, z, reduce_mean=rm, likelihood_weighting=True, t_is=t_or)
        l_or.backward()
        r = abs(l_or.item() - loss.item()) / abs(loss.item())
        rg = ((wo.grad - w.grad).norm() / w.grad.norm()).item()
        print(f"VPSDE({a},{b}) reduce_mean={rm}: t bit-exact; loss rel err {r:.2e}; grad rel err {rg:.2e}")
        assert r < 1e-6 and rg < 1e-5
        cases.append({"a": a, "b": b, "N": 100, "reduce_mean": rm, "t": t_ref, "loss": loss.detach().clone(),
                      "grad_w": w.grad.clone()})
    torch.save({"batch": batch, "u01": u01, "z": z, "w": w0, "cases": cases}, os.path.join(OUT, "dsm_loss_is.pt"))
    print("wrote dsm_loss_is.pt")


if __name__ == "__main__":
    main()
