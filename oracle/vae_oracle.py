"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  fp32 CPU restatement of the unimodal residual encoders /
decoders either side of the score-model path -- `RBlock`, `ResEncoder`, `ResDecoder`, `ResAE.encoder/decoder`,
`ResVAE.encoder/decoder` of h_vae_model_copy.py:9-174 and their CelebA-HQ variants `ResAEN` / `ResVAEN` (:347-588,
family "N": GELU blocks, bilinear up-sampling, sigmoid output) -- in eval mode (BatchNorm uses its running statistics), the
mode every sampling / DSM-training script of the reference runs them in (train_poly_unet_cont.py:556-571: the
autoencoders are loaded from checkpoints and frozen).  SURVEY.md 8f-1: the next row after the score-model path; this
oracle and its golden (tests/golden/res_ae.pt, made by oracle/gen_golden_vae.py from the unmodified reference) are the
parity anchor for it.  State-dict keys are the reference's."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn(sd, p, x, eps=1e-5):
    """nn.BatchNorm2d in eval mode (h_vae_model_copy.py:19,22,47)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=eps)


def rblock(sd, p, x, in_width, out_width, down_rate=None, up_rate=None, family=""):
    """h_vae_model_copy.py:9-39: conv3x3 -> BN -> LeakyReLU(0.2) -> conv3x3 -> BN, 1x1 `size_conv` on the skip when the
    widths differ, LeakyReLU(0.2) AFTER the sum, then average pooling / nearest up-sampling.
    family "N" = `RBlockN` (:347-377): exact GELU instead of LeakyReLU, BILINEAR up-sampling."""
    act = F.gelu if family == "N" else (lambda t: F.leaky_relu(t, 0.2))
    h = F.conv2d(x, sd[p + ".conv.0.weight"], None, padding=1)
    h = act(_bn(sd, p + ".conv.1", h))
    h = _bn(sd, p + ".conv.4", F.conv2d(h, sd[p + ".conv.3.weight"], None, padding=1))
    if in_width != out_width:
        x = F.conv2d(x, sd[p + ".size_conv.weight"], None)
    h = act(x + h)
    if down_rate is not None:
        h = F.avg_pool2d(h, down_rate)
    if up_rate is not None:
        h = F.interpolate(h, scale_factor=up_rate, mode="bilinear" if family == "N" else "nearest")
    return h


def res_encoder(sd, x, channel_list, p="enc", family=""):
    """h_vae_model_copy.py:41-72 -> (mu, logvar); family "N" = `ResEncoderN` (:379-409: LeakyReLU(0.1) stem, RBlockN)."""
    h = F.conv2d(x, sd[p + ".ch_enc.0.weight"], sd[p + ".ch_enc.0.bias"], padding=2)
    h = F.avg_pool2d(F.leaky_relu(_bn(sd, p + ".ch_enc.1", h), 0.1 if family == "N" else 0.2), 2)
    for i, (cin, _mid, cout, rate) in enumerate(channel_list):
        h = rblock(sd, f"{p}.r_blocks.{i}", h, cin, cout, down_rate=rate, family=family)
    mu, logvar = h.chunk(2, dim=1)
    mu = F.linear(mu.reshape(mu.shape[0], -1), sd[p + ".mu_lin.weight"], sd[p + ".mu_lin.bias"])
    logvar = F.linear(logvar.reshape(logvar.shape[0], -1), sd[p + ".logvar_lin.weight"], sd[p + ".logvar_lin.bias"])
    return mu, logvar


def res_decoder(sd, x, channel_list, p="dec", family=""):
    """h_vae_model_copy.py:74-90; family "N" = `ResDecoderN` (:411-428): RBlockN up-blocks, a plain RBlock (LeakyReLU)
    in `ch_dec`, and a Sigmoid after the 5x5 output convolution."""
    h = x
    for i, (cin, _mid, cout, rate) in enumerate(channel_list):
        h = rblock(sd, f"{p}.r_blocks.{i}", h, cin, cout, up_rate=rate, family=family)
    c = channel_list[-1][2]
    h = rblock(sd, f"{p}.ch_dec.0", h, c, c)
    y = F.conv2d(h, sd[p + ".ch_dec.1.weight"], sd[p + ".ch_dec.1.bias"], padding=2)
    return torch.sigmoid(y) if family == "N" else y


def ae_encode(sd, x, enc_channel_list, family=""):
    """ResAE.encoder (h_vae_model_copy.py:164-166) / the mean of ResVAE.encoder (:118-120): the latent the score model
    is trained on (train_poly_unet_cont.py:257-268 stacks these per modality)."""
    return res_encoder(sd, x, enc_channel_list, family=family)[0]


def ae_decode(sd, z, enc_channel_list, dec_channel_list, size_in, family=""):
    """ResAE.decoder / ResVAE.decoder (h_vae_model_copy.py:127-130, 168-171): Linear -> ReLU -> view -> ResDecoder."""
    init = size_in
    for c in enc_channel_list:
        init //= c[3]
    ch = enc_channel_list[-1][2]
    lin = ch * init * init
    side = lin // ch // init
    h = F.relu(F.linear(z, sd["z_lin.weight"], sd["z_lin.bias"]))
    return res_decoder(sd, h.view(z.shape[0], ch, side, side), dec_channel_list, family=family)


def attr_mlp(sd, x, p, n_layers, last_plain=False, eps=1e-5):
    """`enc_net` / `dec_net` of CelebAAttrNewBN[AE] (h_vae_model.py:718-755, 833-868): Linear -> BatchNorm1d (eval) -> ReLU
    blocks at Sequential indices 0, 3, 6, ...; with `last_plain` the final Linear has no norm / activation."""
    h = x
    for i in range(n_layers):
        k = 3 * i
        h = F.linear(h, sd[f"{p}.{k}.weight"], sd[f"{p}.{k}.bias"])
        if last_plain and i == n_layers - 1:
            break
        h = F.batch_norm(h, sd[f"{p}.{k + 1}.running_mean"], sd[f"{p}.{k + 1}.running_var"], sd[f"{p}.{k + 1}.weight"],
                         sd[f"{p}.{k + 1}.bias"], training=False, eps=eps)
        h = F.relu(h)
    return h


def attr_encode(sd, x):
    """CelebAAttrNewBN.encoder (h_vae_model.py:757-760) -> (mu, logvar | None for the AE variant, :870-873)."""
    h = attr_mlp(sd, x, "enc_net", 5)
    mu = F.linear(h, sd["mu_lin.weight"], sd["mu_lin.bias"])
    logvar = F.linear(h, sd["logvar_lin.weight"], sd["logvar_lin.bias"]) if "logvar_lin.weight" in sd else None
    return mu, logvar


def attr_decode(sd, z):
    """CelebAAttrNewBN.decoder (h_vae_model.py:767-768)."""
    return attr_mlp(sd, z, "dec_net", 6, last_plain=True)
