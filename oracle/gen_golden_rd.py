"""Generate tests/golden/rd_predictor.pt: the reverse-diffusion (ancestral) predictor built on the UNMODIFIED
reference's `sde.reverse(score_fn, pf).discretize(x, t)` (/root/reference/sde_helper2.py:236-253, 319-324, 373-381,
465-473) for VPSDE / subVPSDE / VESDE, and check oracle/sde_oracle.py's restatement against it.

Run in the build container only:  python -m oracle.gen_golden_rd
"""
from __future__ import annotations

import os

import torch

from . import sde_oracle as so
from .gen_golden import OUT, import_reference


def main():
    sh, _, _ = import_reference()
    g = torch.Generator().manual_seed(20260)
    cases = []
    for kind, cls, a, b, N in [("vp", sh.VPSDE, 0.1, 20.0, 1000), ("vp", sh.VPSDE, 1.0, 5.0, 100),
                               ("subvp", sh.subVPSDE, 0.1, 20.0, 50), ("ve", sh.VESDE, 0.01, 50.0, 30)]:
        sde = cls(a, b, N)
        spec = so.SdeSpec(kind, a, b, N)
        B = 7
        x = torch.randn(B, 3, 4, 4, generator=g)
        t = torch.rand(B, generator=g) * 0.999 + 1e-3
        t[0] = 1e-3      # index 0 (VE: adjacent sigma = 0)
        t[1] = 1.0       # last index
        score = torch.randn(B, 3, 4, 4, generator=g)
        z = torch.randn(B, 3, 4, 4, generator=g)
        f, G = sde.discretize(x, t)
        fo, Go = so.discretize(spec, x, t)
        assert torch.equal(fo, f) and torch.equal(Go, G), kind
        case = {"kind": kind, "a": a, "b": b, "N": N, "x": x, "t": t, "score": score, "z": z, "disc_f": f, "disc_G": G}
        for pf in (False, True):
            rsde = sde.reverse(lambda xx, tt: score, pf)
            rev_f, rev_G = rsde.discretize(x, t)
            x_mean = x - rev_f
            x_new = x_mean + rev_G[:, None, None, None] * z
            xo, xmo = so.rd_predictor_step(spec, x, t, score, z, pf)
            assert torch.equal(xo, x_new) and torch.equal(xmo, x_mean), (kind, pf)
            case["ode" if pf else "sde"] = {"x": x_new, "x_mean": x_mean}
        cases.append(case)
        print(f"{kind}: oracle == reference (bit-exact), N={N}")
    torch.save(cases, os.path.join(OUT, "rd_predictor.pt"))
    print("wrote rd_predictor.pt")


if __name__ == "__main__":
    main()
