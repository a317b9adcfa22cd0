"""Drop-in for the reference's residual autoencoders `ResAE` / `ResVAE` (h_vae_model_copy.py:9-174) and their CelebA-HQ
variants `ResAEN` / `ResVAEN` (:347-590: GELU blocks, bilinear up-sampling, sigmoid output) -- the encoders
that turn each modality into the latents the score model is trained on, and the decoders that turn sampled latents
back into images (train_poly_unet_cont.py:548-571, 257-268) -- in EVAL mode, the mode every score-model script of the
reference runs them in (loaded from a checkpoint and frozen).  SURVEY.md 8f-1: the first row either side of the path.

Same class names, constructor arguments and `state_dict()` keys as the reference, so its checkpoints load.  Execution:
  * every Conv2d -> BatchNorm2d pair is ONE implicit-GEMM convolution (`sbm_conv_igemm`) with the batch-norm scale
    folded into the bf16 weights and its shift into the bias; the RBlock skip (`x` or the 1x1 `size_conv`) is the GEMM
    epilogue's residual operand;
  * `LeakyReLU(0.2)` after the sum + `AvgPool2d` / `nn.Upsample` = `sbm_lrelu_resample` (one elementwise kernel);
  * the 5x5 input / output convolutions run as im2col rows + a GEMM (`sbm_stem_im2col`), the three Linear layers as
    1x1 GEMMs whose weights are permuted once from the reference's (c, h, w) flattening to channels-last.
Training these nets (`.train()` mode BatchNorm, `reparametrize`) is outside the path: `forward` in train mode raises.
There is no CPU / eager fallback."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L
from . import ops
from .ops import pad8


ACT_LRELU, ACT_GELU = 0, 1
MODE_SAME, MODE_AVGPOOL, MODE_NEAREST, MODE_BILINEAR = 0, 1, 2, 3


def lrelu_resample(x: torch.Tensor, c: int, slope: float, mode: int = 0, rate: int = 1, nchw: bool = False,
                   act: int = ACT_LRELU):
    """act(x) -- LeakyReLU_slope or exact GELU -- then (mode 0) nothing / (1) AvgPool2d(rate) / (2) nearest or
    (3) bilinear up-sampling by rate.  x: channels-last [B,H,W,ld] fp32 or bf16 -> bf16 channels-last, or fp32 NCHW
    when `nchw` (mode 0)."""
    b, h, w, _ = x.shape
    oh, ow = (h // rate, w // rate) if mode == 1 else ((h * rate, w * rate) if mode >= 2 else (h, w))
    out_b = None if nchw else torch.empty((b, oh, ow, pad8(c)), dtype=torch.bfloat16, device=x.device)
    out_n = torch.empty((b, c, h, w), dtype=torch.float32, device=x.device) if nchw else None
    L.check(L.lib().sbm_act_resample(
        L.ptr(x), C.c_int32(L.BF16 if x.dtype == torch.bfloat16 else L.F32), C.c_int64(x.stride(2)), L.ptr(out_b),
        C.c_int64(out_b.stride(2) if out_b is not None else 0), L.ptr(out_n), C.c_int32(b), C.c_int32(h), C.c_int32(w),
        C.c_int32(c), C.c_int32(act), C.c_float(slope), C.c_int32(mode), C.c_int32(max(rate, 1)), L.stream_ptr()),
        "sbm_act_resample")
    return out_n if nchw else out_b


class _Holder(nn.Module):
    """Parameter container with the reference's module structure; executed by the owning model's fused plan."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is executed by ResAE / ResVAE's fused CUDA plan")


class RBlock(_Holder):  # h_vae_model_copy.py:9-39
    _act, _up_mode = ACT_LRELU, MODE_NEAREST      # LeakyReLU(0.2), nn.Upsample(nearest)

    def __init__(self, in_width, middle_width, out_width, down_rate=None, up_rate=None, residual=True):
        super().__init__()
        self.down_rate, self.up_rate, self.residual = down_rate, up_rate, residual
        self.in_width, self.middle_width, self.out_width = in_width, middle_width, out_width
        self.conv = nn.Sequential(
            nn.Conv2d(in_width, middle_width, 3, 1, 1, bias=False), nn.BatchNorm2d(middle_width),
            nn.GELU() if self._act == ACT_GELU else nn.LeakyReLU(0.2),
            nn.Conv2d(middle_width, out_width, 3, 1, 1, bias=False), nn.BatchNorm2d(out_width))
        self.sf = nn.GELU() if self._act == ACT_GELU else nn.LeakyReLU(0.2)
        self.size_conv = nn.Conv2d(in_width, out_width, 1, 1, 0, bias=False)


class RBlockN(RBlock):  # h_vae_model_copy.py:347-377: GELU, bilinear up-sampling
    _act, _up_mode = ACT_GELU, MODE_BILINEAR


class ResEncoder(_Holder):  # h_vae_model_copy.py:41-72
    _block, _stem_slope = RBlock, 0.2

    def __init__(self, channel_list, size_in=64, size_z=64, img_ch=3):
        super().__init__()
        self.img_ch, self.channel_list, self.size_z, self.size_in = img_ch, channel_list, size_z, size_in
        self.ch_enc = nn.Sequential(nn.Conv2d(img_ch, channel_list[0][0], 5, 1, 2), nn.BatchNorm2d(channel_list[0][0]),
                                    nn.LeakyReLU(self._stem_slope), nn.AvgPool2d(2))
        init_size = size_in // 2
        for i in channel_list:
            init_size = init_size // i[3]
        self.final_side = init_size
        self.size_z_lin = (init_size * init_size) * (channel_list[-1][2] // 2)
        self.r_blocks = nn.ModuleList([self._block(*i) for i in channel_list])
        self.mu_lin = nn.Linear(self.size_z_lin, size_z)
        self.logvar_lin = nn.Linear(self.size_z_lin, size_z)


class ResEncoderN(ResEncoder):  # h_vae_model_copy.py:379-409: LeakyReLU(0.1) stem, RBlockN
    _block, _stem_slope = RBlockN, 0.1


class ResDecoder(_Holder):  # h_vae_model_copy.py:74-90
    _block, _sigmoid = RBlock, False

    def __init__(self, channel_list, size_in=64, size_z=64, img_ch=3):
        super().__init__()
        self.img_ch, self.channel_list, self.size_z = img_ch, channel_list, size_z
        self.r_blocks = nn.ModuleList([self._block(i[0], i[1], i[2], None, i[3], True) for i in channel_list])
        c = channel_list[-1][2]
        tail = [RBlock(c, c, c), nn.Conv2d(c, img_ch, 5, 1, 2)] + ([nn.Sigmoid()] if self._sigmoid else [])
        self.ch_dec = nn.Sequential(*tail)


class ResDecoderN(ResDecoder):  # h_vae_model_copy.py:411-428: RBlockN up-blocks, plain RBlock + conv + Sigmoid tail
    _block, _sigmoid = RBlockN, True


class _ResBase(nn.Module):
    _enc_cls, _dec_cls = ResEncoder, ResDecoder

    def __init__(self, enc_channel_list, dec_channel_list, size_in=64, size_z=64, img_ch=3):
        super().__init__()
        self.enc_channel_list, self.dec_channel_list = enc_channel_list, dec_channel_list
        self.size_z, self.size_in, self.img_ch = size_z, size_in, img_ch
        self.enc = self._enc_cls(enc_channel_list, size_in, size_z, img_ch)
        self.dec = self._dec_cls(dec_channel_list, size_in, size_z, img_ch)
        init_size = size_in
        for i in enc_channel_list:
            init_size = init_size // i[3]
        self.size_z_lin = (init_size * init_size) * enc_channel_list[-1][2]
        self.z_lin = nn.Linear(size_z, self.size_z_lin)
        self.z_lin_relu = nn.ReLU()
        self.z_reshape_size = self.size_z_lin // enc_channel_list[-1][2] // init_size
        self._packed: dict = {}

    # ------------------------------------------------------------------ packed operands (re-built when weights change)
    def _cached(self, key, tensors, build):
        sig = tuple((t.data_ptr(), t._version) for t in tensors)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._packed[key] = (sig, val)
        return val

    def _conv_bn(self, conv, bn, im2col=False):
        """Conv2d -> BatchNorm2d (eval) as one GEMM operand: W' = W * s[o], b' = (bias - mean) * s + beta,
        s = gamma / sqrt(var + eps)."""
        deps = (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var) + \
            ((conv.bias,) if conv.bias is not None else ())

        def build():
            s = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
            w = conv.weight.float() * s[:, None, None, None]
            b0 = conv.bias.float() if conv.bias is not None else torch.zeros_like(s)
            bias = ((b0 - bn.running_mean.float()) * s + bn.bias.float()).contiguous()
            wpk = ops.pack_linear_weight(w.reshape(w.shape[0], -1)) if im2col else ops.pack_conv2d_weight(w)
            return wpk, bias

        return self._cached((id(conv), "bn"), deps, build)

    def _conv_plain(self, conv, im2col=False):
        def build():
            w = conv.weight.float()
            return ops.pack_linear_weight(w.reshape(w.shape[0], -1)) if im2col else ops.pack_conv2d_weight(w)
        return self._cached((id(conv), "plain"), (conv.weight,), build)

    def _lin_from_chw(self, lin, ch, side):
        """nn.Linear over a (c, h, w)-flattened map -> operand over the channels-last (h, w, c) flattening."""
        def build():
            w = lin.weight.float().reshape(lin.out_features, ch, side, side).permute(0, 2, 3, 1)
            return ops.pack_linear_weight(w.reshape(lin.out_features, -1).contiguous())
        return self._cached((id(lin), "chw_in"), (lin.weight,), build)

    def _lin_to_chw(self, lin, ch, side):
        """nn.Linear whose output is viewed as (c, h, w) -> rows re-ordered so the GEMM writes (h, w, c)."""
        def build():
            w = lin.weight.float().reshape(ch, side, side, lin.in_features).permute(1, 2, 0, 3)
            b = lin.bias.float().reshape(ch, side, side).permute(1, 2, 0)
            return ops.pack_linear_weight(w.reshape(-1, lin.in_features).contiguous()), b.reshape(-1).contiguous()
        return self._cached((id(lin), "chw_out"), (lin.weight, lin.bias), build)

    # ------------------------------------------------------------------ fused plan
    def _check(self, t):
        if self.training:
            raise NotImplementedError("the B200 path runs the residual autoencoders in eval() mode only (frozen, as in "
                                      "the reference's score-model scripts); training them is outside the path")
        if not t.is_cuda:
            raise L.SbmError("ResAE / ResVAE need CUDA tensors: the B200 path has no CPU fallback")

    def _rblock(self, blk: RBlock, x_b):
        """x_b: bf16 channels-last -> bf16 channels-last (after the block's pooling / up-sampling)."""
        c_in, c_mid, c_out = blk.in_width, blk.middle_width, blk.out_width
        w1, b1 = self._conv_bn(blk.conv[0], blk.conv[1])
        act, up = blk._act, blk._up_mode
        h = ops.conv_igemm(x_b, w1, kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_mid, bias=b1)
        a = lrelu_resample(h, c_mid, 0.2, act=act)
        if c_in != c_out:
            res = ops.conv_igemm(x_b, self._conv_plain(blk.size_conv), kind=L.CONV_S1, kh=1, kw=1, cin=c_in, cout=c_out)
        else:
            res = x_b
        w2, b2 = self._conv_bn(blk.conv[3], blk.conv[4])
        h2 = ops.conv_igemm(a, w2, kind=L.CONV_S1, kh=3, kw=3, cin=c_mid, cout=c_out, bias=b2, residual=res)
        if blk.down_rate is not None:
            return lrelu_resample(h2, c_out, 0.2, MODE_AVGPOOL, blk.down_rate, act=act)
        if blk.up_rate is not None:
            return lrelu_resample(h2, c_out, 0.2, up, blk.up_rate, act=act)
        return lrelu_resample(h2, c_out, 0.2, act=act)

    @torch.no_grad()
    def _encode(self, x):
        self._check(x)
        enc = self.enc
        b = x.shape[0]
        x = x.contiguous().float()
        c0 = enc.channel_list[0][0]
        w0, b0 = self._conv_bn(enc.ch_enc[0], enc.ch_enc[1], im2col=True)
        a0 = ops.stem_im2col(x, 5, 5)
        h = ops.conv_igemm(a0, w0, kind=L.CONV_S1, kh=1, kw=1, cin=self.img_ch * 25, cout=c0, bias=b0)
        cur = lrelu_resample(h, c0, enc._stem_slope, MODE_AVGPOOL, 2)
        for blk in enc.r_blocks:
            cur = self._rblock(blk, cur)
        ch = enc.channel_list[-1][2]
        half, side = ch // 2, enc.final_side
        outs = []
        for lin, lo in ((enc.mu_lin, 0), (enc.logvar_lin, half)):
            flat = cur[..., lo:lo + half].contiguous().view(b, 1, 1, side * side * half)
            y = ops.conv_igemm(flat, self._lin_from_chw(lin, half, side), kind=L.CONV_S1, kh=1, kw=1,
                               cin=side * side * half, cout=self.size_z, bias=lin.bias)
            outs.append(y.view(b, -1)[:, :self.size_z].contiguous())
        return outs[0], outs[1]

    @torch.no_grad()
    def decoder(self, z):
        self._check(z)
        b = z.shape[0]
        ch = self.enc_channel_list[-1][2]
        side = self.z_reshape_size
        zb, _ = ops.nchw_to_nhwc(z.contiguous().float().view(b, self.size_z, 1, 1))
        wz, bz = self._lin_to_chw(self.z_lin, ch, side)
        h = ops.conv_igemm(zb, wz, kind=L.CONV_S1, kh=1, kw=1, cin=self.size_z, cout=self.size_z_lin, bias=bz)
        cur = lrelu_resample(h, self.size_z_lin, 0.0).view(b, side, side, ch)       # ReLU; rows are (h, w, c)
        for blk in self.dec.r_blocks:
            cur = self._rblock(blk, cur)
        last = self.dec.ch_dec[0]
        c = last.out_width
        # ch_dec: RBlock without resampling, then the 5x5 output convolution as im2col rows + GEMM
        w1, b1 = self._conv_bn(last.conv[0], last.conv[1])
        h1 = ops.conv_igemm(cur, w1, kind=L.CONV_S1, kh=3, kw=3, cin=c, cout=c, bias=b1)
        a1 = lrelu_resample(h1, c, 0.2)
        w2, b2 = self._conv_bn(last.conv[3], last.conv[4])
        h2 = ops.conv_igemm(a1, w2, kind=L.CONV_S1, kh=3, kw=3, cin=c, cout=c, bias=b2, residual=cur)
        feat = lrelu_resample(h2, c, 0.2, nchw=True)
        oc = self.dec.ch_dec[1]
        cols = ops.stem_im2col(feat, 5, 5)
        y = ops.conv_igemm(cols, self._conv_plain(oc, im2col=True), kind=L.CONV_S1, kh=1, kw=1, cin=c * 25,
                           cout=self.img_ch, bias=oc.bias, nchw=True)
        return torch.sigmoid_(y) if self.dec._sigmoid else y       # ResDecoderN's nn.Sigmoid on the [B, img_ch, H, W] image

    def sample(self, amount, device):
        return self.decoder(torch.randn(amount, self.size_z).to(device))


class ResAE(_ResBase):  # h_vae_model_copy.py:145-174
    def encoder(self, x):
        return self._encode(x)[0]

    def forward(self, m):
        return self.decoder(self.encoder(m))


class ResAEN(ResAE):  # h_vae_model_copy.py:549-590 (CelebA-HQ image modality, train_lat_celebhq_unet_cont2.py:427-431)
    _enc_cls, _dec_cls = ResEncoderN, ResDecoderN


class ResVAE(_ResBase):  # h_vae_model_copy.py:92-143
    def encoder(self, x):
        return self._encode(x)

    def reparametrize(self, mu, logvar):
        noise = torch.normal(mean=0, std=1, size=mu.shape).to(mu.device)   # CPU draw + copy, like the reference (:123)
        return mu + torch.exp(logvar / 2) * noise

    def forward(self, m):
        mu, logvar = self.encoder(m)
        z = self.reparametrize(mu, logvar)
        return self.decoder(z), mu, logvar


class ResVAEN(ResVAE):  # h_vae_model_copy.py:457-503
    _enc_cls, _dec_cls = ResEncoderN, ResDecoderN
