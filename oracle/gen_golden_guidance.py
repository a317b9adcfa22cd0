"""Generate tests/golden/guidance.pt: the UNMODIFIED reference's em_predictor / corrector (/root/reference/sde_helper2.py
:45-106, 277-317) with classifier / EBM guidance ON (cl_g = three pair energy nets, cl_s), and check the oracle's
restatement (oracle/guidance_oracle.py + oracle/sde_oracle.py) against them.

Run in the build container only:  python -m oracle.gen_golden_guidance
"""
from __future__ import annotations

import os

import torch

from . import guidance_oracle as go
from . import sde_oracle as so
from .det_weights import fill_state_dict
from .gen_golden import OUT, NoiseFeed, import_reference

SIZE_Z, HIDDEN, TIME_DIM = 256, 128, 64


def shapes(index_conditioned=False):
    s = {"time_proj.weight": (HIDDEN, TIME_DIM), "time_proj.bias": (HIDDEN,), "fc1.weight": (HIDDEN, 2 * SIZE_Z),
         "fc1.bias": (HIDDEN,), "fc2.weight": (HIDDEN, HIDDEN), "fc2.bias": (HIDDEN,), "fc3.weight": (1, HIDDEN),
         "fc3.bias": (1,)}
    if index_conditioned:
        s.update({"id_emb1.weight": (16, HIDDEN), "id_emb2.weight": (16, HIDDEN)})
    return s


def pair_weights(pair):
    return fill_state_dict({f"{pair}.{k}": v for k, v in shapes().items()}, gain=2.0)


def main():
    sh, _, _ = import_reference()
    g = torch.Generator().manual_seed(606)
    B, M, D = 6, 3, 16
    sde = sh.VPSDE(0.1, 20.0, 1000)
    spec = so.SdeSpec("vp", 0.1, 20.0, 1000)
    x = torch.randn(B, M, D, D, generator=g)
    t = torch.rand(B, generator=g) * 0.9 + 0.05
    score = torch.randn(B, M, D, D, generator=g)
    z_pred = torch.randn(B, M, D, D, generator=g)
    z_corr = torch.randn(B, M, D, D, generator=g)
    sds = {}
    cl_g = {}
    for pair in ("01", "02", "12"):
        sd = {k.split(".", 1)[1]: v for k, v in pair_weights(pair).items()}
        sds[pair] = sd
        cl_g[pair] = (lambda sd: (lambda flat, tt: go.energy(sd, flat, tt, time_dim=TIME_DIM)))(sd)
    cl_s = 30.0     # makes the guidance term about as large as the score itself (the energy gradient carries the 1/B
    # of cl_out.mean(); the reference sweeps cl_s from 1 to 5e4, train_lat_celebhq_unet_cont2.py:580-582)
    cases = []
    for given in ("0", "12", "1"):
        score_fn = lambda xx, tt: score.clone()       # the reference edits the returned tensor in place (:75)
        with NoiseFeed([z_pred]).patched():
            xp, xpm = sh.em_predictor(x, t, score_fn, sde, cl_g=cl_g, cl_s=cl_s, given=given, all_mods="012")
        with NoiseFeed([z_corr]).patched():
            xc, xcm = sh.corrector(x, t, score_fn, sde, 1, 0.16, cl_g=cl_g, cl_s=cl_s, given=given, all_mods="012")
        gs = go.guided_score(score, x, t, cl_g, cl_s, given, "012")
        op, opm = so.em_predictor_step(spec, x, t, gs, z_pred)
        oc, ocm = so.corrector_step(spec, x, t, gs, z_corr, 0.16)
        assert torch.equal(op, xp) and torch.equal(opm, xpm) and torch.equal(oc, xc) and torch.equal(ocm, xcm), given
        share = ((gs - score).norm() / score.norm()).item()
        print(f"given={given!r}: oracle == reference (bit-exact); guidance changes the score by {share:.2f} of its norm")
        assert 0.3 < share < 3.0
        cases.append({"given": given, "guided_score": gs, "pred_x": xp, "pred_mean": xpm, "corr_x": xc,
                      "corr_mean": xcm})
    torch.save({"x": x, "t": t, "score": score, "z_pred": z_pred, "z_corr": z_corr, "cl_s": cl_s, "sde": (0.1, 20.0, 1000),
                "size_z": SIZE_Z, "hidden": HIDDEN, "time_dim": TIME_DIM, "cases": cases},
               os.path.join(OUT, "guidance.pt"))
    print("wrote guidance.pt")


if __name__ == "__main__":
    main()
