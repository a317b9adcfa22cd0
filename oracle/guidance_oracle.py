"""ORACLE (test infrastructure, NOT product code): the guidance branch of the reference's samplers.

`guided_score` restates /root/reference/sde_helper2.py:65-94 (corrector) and :283-312 (RSDE.sde): for every pair in
('01','02','12') joining an observed and a predicted modality, score[:, m] -= cl_s * d mean(E(new_x, t)) / d new_x.
`energy` is the fp32 functional form of this package's ClwithTime2 / ClwithTime3 specification (the upstream classes are
absent from the reference repository; see score_based_multimodal_autoencoder_b200/guidance.py) over a plain state dict.
Pinned by oracle/gen_golden_guidance.py: the UNMODIFIED reference's em_predictor / corrector run with these energy nets
as `cl_g` reproduce `guided_score` + the oracle step functions bit for bit.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def energy(sd, x, t, id1=None, id2=None, time_dim=64):
    half = time_dim // 2
    freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * freqs[None, :]
    temb = torch.cat((arg.sin(), arg.cos()), dim=-1)
    e = F.linear(temb, sd["time_proj.weight"], sd["time_proj.bias"])
    if id1 is not None:
        e = e + sd["id_emb1.weight"][int(id1)] + sd["id_emb2.weight"][int(id2)]
    h1 = F.silu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]) + e)
    h2 = F.silu(F.linear(h1, sd["fc2.weight"], sd["fc2.bias"]))
    return F.linear(h2, sd["fc3.weight"], sd["fc3.bias"])


def energy_grad(net, new_x, t, *ids):
    with torch.enable_grad():
        nx = new_x.detach().clone().requires_grad_(True)
        out = net(nx.view(nx.shape[0], -1), t, *ids)
        return torch.autograd.grad(out.mean(), nx)[0]


def guided_score(score, x, t, cl_g, cl_s, given, all_mods):
    """sde_helper2.py:65-94 / 283-312; cl_g: dict pair -> callable(flat, t)."""
    score = score.clone()
    if cl_g is None or not given:
        return score
    predicted = "".join(m for m in all_mods if m not in given)
    base = int(all_mods[0])
    for a, b in (("0", "1"), ("0", "2"), ("1", "2")):
        if (a in given and b in predicted) or (b in given and a in predicted):
            m1, m2 = int(a) - base, int(b) - base
            new_x = torch.cat([x[:, m1].unsqueeze(1), x[:, m2].unsqueeze(1)], dim=1)
            g = energy_grad(cl_g[a + b], new_x, t)
            score[:, m1] -= cl_s * g[:, 0]
            score[:, m2] -= cl_s * g[:, 1]
    return score
