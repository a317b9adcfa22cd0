"""Classifier / energy-based guidance of the score inside the samplers.

The reference's `corrector` and `RSDE.sde` (sde_helper2.py:65-94, 283-312) differentiate a pairwise energy net w.r.t.
the two modality latents it sees and do `score[:, m] -= cl_s * grad`:

    new_x   = cat(x[:, m1], x[:, m2]).view(B, 2 * size_z)            # m = int(mod) - int(all_mods[0])
    cl_out  = cl_g[pair](new_x, t)                                    # [B, n_class]
    grad    = autograd.grad(cl_out.mean(), new_x)
    score[:, m1] -= cl_s * grad[:, 0];  score[:, m2] -= cl_s * grad[:, 1]

for every pair in ('01', '02', '12') that joins an observed and a predicted modality.  The energy-net classes the
scripts import (`lat_sm2_model.ClwithTime2 / ClwithTime3`) are NOT in the reference repository; their contracts are
fixed by the call sites:

  * `ClwithTime2(n_mod=2, size_z, n_class=1)`: `net(flat[B, n_mod*size_z], t[B]) -> [B, n_class]`
    (train_cel_clwithtime_ebm_NOIND.py:145-150, 318; one net per pair: train_lat_celebhq_unet_cont2.py:487);
  * `ClwithTime3(n_mod=2, size_z, n_class=1)`: `net(flat, t, id1, id2) -> [B, n_class]` with the two modality indices
    (train_poly_clwithtime_ebm_IND.py:135, 259; sampler call train_poly_unet_cont.py:72-88, which draws ONE random
    (observed, predicted) pair per score evaluation and updates the predicted modality only).

Here both are time-conditioned two-hidden-layer MLPs (the architecture below is this package's own, documented choice):

    e   = Linear(time_dim -> hidden)(sinusoidal(t, time_dim))   [+ Embedding(id1) + Embedding(id2) for ClwithTime3]
    h1  = SiLU(Linear(n_mod*size_z -> hidden)(x) + e);  h2 = SiLU(Linear(hidden -> hidden)(h1))
    out = Linear(hidden -> n_class)(h2)

`forward` is plain torch (training the energy nets is outside the accelerated path); the samplers call `energy_grad`,
which evaluates d mean(out) / d x with the library's kernels: the pair gather (`sbm_guidance_gather`), `sbm_time_embed`,
five tcgen05 GEMMs (`sbm_conv_igemm`: the time projection, two forward layers keeping their pre-activations, two
transposed layers for the input gradient), two `sbm_act_bwd` and the in-place score update (`sbm_guidance_apply`).
Any other callable passed as `cl_g[pair]` takes the reference's own route (torch.autograd through the callable).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _lib as L
from . import ops
from .ops import pad8


class ClwithTime2(nn.Module):
    """Pairwise energy / classifier net over the concatenated latents of `n_mod` modalities, conditioned on t."""

    def __init__(self, n_mod=2, size_z=64, n_class=1, hidden=512, time_dim=128):
        super().__init__()
        self.n_mod, self.size_z, self.n_class, self.hidden, self.time_dim = n_mod, size_z, n_class, hidden, time_dim
        self.time_proj = nn.Linear(time_dim, hidden)
        self.fc1 = nn.Linear(n_mod * size_z, hidden)
        self.fc2 = nn.Linear(hidden, hidden)
        self.fc3 = nn.Linear(hidden, n_class)
        self._packed: dict = {}

    # ---- torch forward (training / reference semantics)
    def _temb(self, t):
        half = self.time_dim // 2
        freqs = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
        arg = t[:, None] * freqs[None, :]
        return torch.cat((arg.sin(), arg.cos()), dim=-1)      # unet_model.py:40-47 layout (sbm_time_embed mode 0)

    def _cond(self, t, id1=None, id2=None):
        return self.time_proj(self._temb(t.float()))

    def forward(self, x, t, id1=None, id2=None):
        e = self._cond(t, id1, id2)
        h1 = nn.functional.silu(self.fc1(x) + e)
        h2 = nn.functional.silu(self.fc2(h1))
        return self.fc3(h2)

    # ---- kernel path: d mean(out) / d x
    def _cached(self, key, params, build):
        sig = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._packed[key] = (sig, val)
        return val

    def _w(self, lin):      # forward operand  [1, O, pad8(I)]
        return self._cached((id(lin), "fw"), (lin.weight,), lambda: ops.pack_linear_weight(lin.weight))

    def _wt(self, lin):     # transposed operand for the input gradient: rows = I, cols = O
        w = lin.weight
        o, i = w.shape
        return self._cached((id(lin), "bw"), (w,), lambda: ops.pack_weight(w.detach().contiguous(), 1, i, o, 0, 1, i))

    def _extra_rowbias(self, e, id1, id2):
        return e

    @torch.no_grad()
    def energy_grad_rows(self, xb, t, id1=None, id2=None):
        """xb: bf16 [B,1,1,pad8(n_mod*size_z)] rows of the gathered latents -> fp32 [B,1,1,pad8(n_mod*size_z)] =
        d mean(net(x, t)) / d x  (mean over batch AND classes, like `cl_out.mean()`)."""
        b = xb.shape[0]
        din, hid = self.n_mod * self.size_z, self.hidden
        te = ops.time_embed(t.contiguous().float(), self.time_dim, 0).view(b, 1, 1, -1)
        e = ops.conv_igemm(te, self._w(self.time_proj), kind=L.CONV_S1, kh=1, kw=1, cin=self.time_dim, cout=hid,
                           bias=self.time_proj.bias)
        e = self._extra_rowbias(e, id1, id2)
        a1 = torch.empty((b, 1, 1, pad8(hid)), dtype=torch.bfloat16, device=xb.device)     # pre-activation of layer 1
        h1 = ops.conv_igemm(xb, self._w(self.fc1), kind=L.CONV_S1, kh=1, kw=1, cin=din, cout=hid, bias=self.fc1.bias,
                            rowbias=e.view(b, -1), act=L.ACT_SILU, out_dtype=torch.bfloat16, out2=a1, out2_preact=True)
        a2 = ops.conv_igemm(h1, self._w(self.fc2), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=hid, bias=self.fc2.bias)
        # d mean(out) / d h2 = column sums of fc3.weight / (B * n_class): the same row for every sample
        g3 = self._cached("g3", (self.fc3.weight,), lambda: torch.zeros(pad8(hid), device=xb.device).index_add_(
            0, torch.arange(hid, device=xb.device), self.fc3.weight.detach().float().sum(0)))
        # rows = samples along dim 2 (the row stride the kernels read is stride(2)): the broadcast row has stride 0
        dy = (g3 / float(b * self.n_class)).view(1, 1, 1, -1).expand(1, 1, b, -1)
        _, da2 = ops.act_bwd(dy, a2.view(1, 1, b, -1), hid, L.ACT_SILU)
        dh1 = ops.conv_igemm(da2.view(b, 1, 1, -1), self._wt(self.fc2), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=hid)
        _, da1 = ops.act_bwd(dh1.view(1, 1, b, -1), a1.view(1, 1, b, -1), hid, L.ACT_SILU)
        return ops.conv_igemm(da1.view(b, 1, 1, -1), self._wt(self.fc1), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=din)

    def energy_grad(self, x_flat, t, id1=None, id2=None):
        """d mean(self(x, t)) / d x for x_flat [B, n_mod*size_z] (fp32 CUDA) -> [B, n_mod*size_z] fp32."""
        if not x_flat.is_cuda:
            raise L.SbmError("energy_grad needs CUDA tensors (no CPU fallback)")
        b, din = x_flat.shape
        xb = torch.zeros((b, 1, 1, pad8(din)), dtype=torch.bfloat16, device=x_flat.device)
        xb[..., :din] = x_flat.view(b, 1, 1, din)
        return self.energy_grad_rows(xb, t, id1, id2).view(b, -1)[:, :din].contiguous()


class ClwithTime3(ClwithTime2):
    """Index-conditioned variant: ONE net for every modality pair, told which two modalities it sees."""

    def __init__(self, n_mod=2, size_z=64, n_class=1, hidden=512, time_dim=128, max_mods=16):
        super().__init__(n_mod, size_z, n_class, hidden, time_dim)
        self.id_emb1 = nn.Embedding(max_mods, hidden)
        self.id_emb2 = nn.Embedding(max_mods, hidden)

    def _cond(self, t, id1=None, id2=None):
        e = super()._cond(t)
        return e + self.id_emb1.weight[int(id1)] + self.id_emb2.weight[int(id2)]

    def _extra_rowbias(self, e, id1, id2):
        add = (self.id_emb1.weight[int(id1)] + self.id_emb2.weight[int(id2)]).detach().float()
        b = e.shape[0]
        row = torch.zeros(e.shape[-1], device=e.device)
        row[:self.hidden] = add
        ev = e.view(1, 1, b, -1)                                  # rows along dim 2, broadcast row with stride 0
        ops.add(ev, row.view(1, 1, 1, -1).expand(1, 1, b, -1), self.hidden, out=ev)
        return e


def _pairs(given, all_mods, cl_g):
    """The reference's three hard-wired pairs (sde_helper2.py:68, 79, 88), generalised to every key of `cl_g` that joins
    an observed and a predicted modality.  Yields (key, channel of mod1, channel of mod2)."""
    predicted = "".join(m for m in all_mods if m not in given)
    base = int(all_mods[0])
    for key in cl_g:
        if len(key) != 2:
            continue
        a, b = key[0], key[1]
        if (a in given and b in predicted) or (b in given and a in predicted):
            yield key, int(a) - base, int(b) - base


def _autograd_grad(net, new_x, t, *ids):
    """The reference's own route for an arbitrary callable: torch.autograd through `net`."""
    with torch.enable_grad():
        nx = new_x.detach().clone().requires_grad_(True)
        out = net(nx.view(nx.shape[0], -1), t, *ids)
        return torch.autograd.grad(out.mean(), nx)[0]


def apply_guidance(score, x, t, cl_g, cl_s, given, all_mods):
    """score (fp32 [B,M,D,D], updated IN PLACE like the reference and returned) -= cl_s * d mean(E) / d x on the two
    channels of every (observed, predicted) pair."""
    if not score.is_cuda:
        raise L.SbmError("guidance runs on CUDA tensors (no CPU fallback)")
    score = score if score.is_contiguous() else score.contiguous()
    x = x.detach().contiguous().float()
    b, m, d1, d2 = x.shape
    dd = d1 * d2
    ls = L.LatentShape(b, m, dd)
    if isinstance(cl_g, dict):
        todo = [(cl_g[k], m1, m2, m1, m2, ()) for k, m1, m2 in _pairs(given, all_mods, cl_g)]
    else:
        # index-conditioned net (train_poly_unet_cont.py:72-88): one random (observed, predicted) pair per call, only
        # the predicted modality is updated
        if cl_s is None:
            return score
        predicted = "".join(k for k in all_mods if k not in given)
        base = int(all_mods[0])
        mod1 = given[torch.randint(len(given), (1,)).item()]
        mod2 = predicted[torch.randint(len(predicted), (1,)).item()]
        i1, i2 = int(mod1) - base, int(mod2) - base
        todo = [(cl_g, i1, i2, -1, i2, (i1, i2))]
    for net, m1, m2, u1, u2, ids in todo:
        if isinstance(net, ClwithTime2):
            ld = pad8(2 * dd)
            xb = torch.empty((b, 1, 1, ld), dtype=torch.bfloat16, device=x.device)
            L.check(L.lib().sbm_guidance_gather(C.byref(ls), L.ptr(x), C.c_int32(m1), C.c_int32(m2), L.ptr(xb),
                                                C.c_int32(ld), L.stream_ptr()), "sbm_guidance_gather")
            g = net.energy_grad_rows(xb, t, *ids)
            ldg = g.stride(0)
        else:
            new_x = torch.stack((x[:, m1], x[:, m2]), dim=1)
            g = _autograd_grad(net, new_x, t, *ids).reshape(b, 2 * dd).contiguous().float()
            ldg = 2 * dd
        L.check(L.lib().sbm_guidance_apply(C.byref(ls), L.ptr(score), L.ptr(g), C.c_int64(ldg), C.c_int32(u1),
                                           C.c_int32(u2), C.c_float(float(cl_s)), L.stream_ptr()), "sbm_guidance_apply")
    return score
