"""Drop-in for the attribute-modality autoencoders of the reference, `CelebAAttrNewBN` / `CelebAAttrNewBNAE`
(h_vae_model.py:712-775, 828-899; train_lat_celebhq_unet_cont2.py:459-461), in EVAL mode (frozen, loaded from a
checkpoint, as in every score-model script).  SURVEY.md 8f-1.

Same class names, constructor arguments and `state_dict()` keys.  Every `Linear -> BatchNorm1d -> ReLU` block is ONE
GEMM on `sbm_conv_igemm` (1x1, batch-norm scale folded into the bf16 weights, shift into the bias) followed by the ReLU
of `sbm_act_resample` (slope 0), which also writes the bf16 operand of the next GEMM.  No CPU / eager fallback; train()
mode raises (training these nets stays with the reference)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from .h_vae_model_copy import lrelu_resample


def _mlp(widths, last_plain=False):
    layers = []
    for i in range(len(widths) - 1):
        layers.append(nn.Linear(widths[i], widths[i + 1]))
        if not (last_plain and i == len(widths) - 2):
            layers += [nn.BatchNorm1d(widths[i + 1]), nn.ReLU()]
    return nn.Sequential(*layers)


class _AttrBase(nn.Module):
    def __init__(self, size_z, att_size, with_logvar):
        super().__init__()
        self.size_z, self.att_size = size_z, att_size
        self.enc_net = _mlp([att_size, 128, 256, 512, 512, 512])
        self.mu_lin = nn.Linear(512, size_z)
        if with_logvar:
            self.logvar_lin = nn.Linear(512, size_z)
        self.dec_net = _mlp([size_z, 512, 512, 512, 256, 128, att_size], last_plain=True)
        self._packed: dict = {}

    def _cached(self, key, tensors, build):
        sig = tuple((t.data_ptr(), t._version) for t in tensors)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._packed[key] = (sig, val)
        return val

    def _lin_bn(self, lin, bn):
        """Linear -> BatchNorm1d (eval): W' = W * s[:, None], b' = (b - mean) * s + beta, s = gamma / sqrt(var + eps)."""
        def build():
            s = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
            return (ops.pack_linear_weight((lin.weight.float() * s[:, None]).contiguous()),
                    ((lin.bias.float() - bn.running_mean.float()) * s + bn.bias.float()).contiguous())
        return self._cached((id(lin), "bn"), (lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var),
                            build)

    def _lin(self, lin):
        return self._cached((id(lin), "plain"), (lin.weight,), lambda: ops.pack_linear_weight(lin.weight.float()))

    def _check(self, t):
        if self.training:
            raise NotImplementedError("the B200 path runs the attribute autoencoders in eval() mode only (frozen, as in "
                                      "the reference's score-model scripts)")
        if not t.is_cuda:
            raise L.SbmError("CelebAAttrNewBN[AE] need CUDA tensors: the B200 path has no CPU fallback")

    def _run(self, net, x):
        """x: fp32 [B, C] -> fp32 [B, C_out] (plain last layer) or bf16 channels-last [B,1,1,C_out] (after a ReLU)."""
        b = x.shape[0]
        cur, _ = ops.nchw_to_nhwc(x.contiguous().float().view(b, x.shape[1], 1, 1))
        mods = list(net)
        i = 0
        while i < len(mods):
            lin = mods[i]
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
                w, bias = self._lin_bn(lin, mods[i + 1])
                h = ops.conv_igemm(cur, w, kind=L.CONV_S1, kh=1, kw=1, cin=lin.in_features, cout=lin.out_features,
                                   bias=bias)
                cur = lrelu_resample(h, lin.out_features, 0.0)          # ReLU, bf16 operand of the next GEMM
                i += 3
            else:
                h = ops.conv_igemm(cur, self._lin(lin), kind=L.CONV_S1, kh=1, kw=1, cin=lin.in_features,
                                   cout=lin.out_features, bias=lin.bias)
                return h.view(b, -1)[:, :lin.out_features].contiguous()
        return cur

    def _head(self, feat, lin):
        b = feat.shape[0]
        y = ops.conv_igemm(feat, self._lin(lin), kind=L.CONV_S1, kh=1, kw=1, cin=lin.in_features, cout=lin.out_features,
                           bias=lin.bias)
        return y.view(b, -1)[:, :lin.out_features].contiguous()

    @torch.no_grad()
    def decoder(self, z):
        self._check(z)
        return self._run(self.dec_net, z)

    def sample(self, amount, device):
        return self.decoder(torch.randn(amount, self.size_z).to(device))


class CelebAAttrNewBN(_AttrBase):  # h_vae_model.py:712-775
    def __init__(self, size_z=64, att_size=18):
        super().__init__(size_z, att_size, with_logvar=True)

    @torch.no_grad()
    def encoder(self, x):
        self._check(x)
        feat = self._run(self.enc_net, x)
        return self._head(feat, self.mu_lin), self._head(feat, self.logvar_lin)

    def reparametrize(self, mu, logvar):
        noise = torch.normal(mean=0, std=1, size=mu.shape).to(mu.device)
        return mu + torch.exp(logvar / 2) * noise

    def forward(self, m):
        mu, logvar = self.encoder(m)
        return self.decoder(self.reparametrize(mu, logvar)), mu, logvar


class CelebAAttrNewBNAE(_AttrBase):  # h_vae_model.py:828-899
    def __init__(self, size_z=64):
        super().__init__(size_z, 18, with_logvar=False)

    @torch.no_grad()
    def encoder(self, x):
        self._check(x)
        return self._head(self._run(self.enc_net, x), self.mu_lin)

    def forward(self, m):
        return self.decoder(self.encoder(m))
