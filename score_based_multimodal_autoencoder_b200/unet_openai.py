"""Drop-in for the reference's `unet_openai.py` score network (`UNetModel`, unet_openai.py:361-575).

Same constructor signature, same parameter names / shapes (`state_dict()` of the reference loads unchanged,
SURVEY.md Appendix D) and the same `forward(x, timesteps, z=None, y=None)` contract.  The sub-modules are parameter
containers mirroring the reference's module tree; the arithmetic is issued by `UNetModel.forward` as hand-written
sm_100a kernels through the C ABI (ops.py -> libsbmae_b200.so):

  * every convolution / linear = the tcgen05 implicit-GEMM kernel (3x3, 3x3 stride 2, 1x1, linear), with the
    `h + emb_out[..., None, None]` of unet_openai.py:303 fused as a per-sample row bias and `skip(x) + h`
    (unet_openai.py:305) as the residual of the second convolution's epilogue;
  * GroupNorm32 (unet_openai.py:10-12, statistics in fp64) + SiLU = group-statistics + apply kernels;
  * `QKVAttention` (unet_openai.py:345-358) = the softmax-attention core kernel with the per-head [q|k|v] layout;
  * all per-block `emb_layers` projections (unet_openai.py:255-261) are ONE GEMM per forward; the time MLP and the
    z projection (unet_openai.py:421-433, 553-559) share their second GEMM (concatenated K);
  * `torch.cat([h, hs.pop()])` (unet_openai.py:571) never materialises: producers write into the two channel ranges
    of a pre-planned concat buffer.

With autograd enabled the forward runs through `autograd_openai.py` (one autograd node, hand-written backward): the
z-conditioned DSM training of `train_lat_celebhq_unet_cont2_cond.py`; train-mode dropout = `sbm_dropout` (Philox masks).
There is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _lib as L
from . import ops
from .ops import pad8


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is executed by UNetModel.forward's fused CUDA plan")


class GroupNorm32(nn.GroupNorm):  # unet_openai.py:10-12
    pass


def conv_nd(dims, *args, **kwargs):  # unet_openai.py:15-25
    if dims == 1:
        return nn.Conv1d(*args, **kwargs)
    if dims == 2:
        return nn.Conv2d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims} (the latent score nets are 2-D)")


def linear(*args, **kwargs):  # unet_openai.py:28-32
    return nn.Linear(*args, **kwargs)


def zero_module(module):  # unet_openai.py:45-51
    for p in module.parameters():
        p.detach().zero_()
    return module


def normalization(channels):  # unet_openai.py:54-63
    return GroupNorm32(32, channels)


def timestep_embedding(timesteps, dim, max_period=10000):
    """unet_openai.py:66-83 on the device: [cos | sin] sinusoid of `timesteps` (fp32 [B, dim])."""
    b = timesteps.shape[0]
    ld = pad8(dim)
    out_b = torch.empty((b, ld), dtype=torch.bfloat16, device=timesteps.device)
    out_f = torch.zeros((b, ld), dtype=torch.float32, device=timesteps.device)
    if max_period != 10000:
        raise NotImplementedError("max_period is fixed at 10000 (the reference never passes another value)")
    import ctypes as C
    L.check(L.lib().sbm_time_embed(L.ptr(timesteps.contiguous().float()), L.ptr(out_b), L.ptr(out_f), C.c_int32(b),
                                   C.c_int32(dim), C.c_int32(ld), C.c_int32(1), L.stream_ptr()), "sbm_time_embed")
    return out_f[:, :dim]


class TimestepEmbedSequential(nn.Sequential):  # unet_openai.py:144-158
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("TimestepEmbedSequential is executed by UNetModel.forward's fused CUDA plan")


class Upsample(_Container):  # unet_openai.py:161-188
    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, channels, channels, 3, padding=1)


class Downsample(_Container):  # unet_openai.py:191-213
    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        if not use_conv:
            # unet_openai.py:209 calls `avg_pool_nd(stride)`: the stride lands in the `dims` parameter and
            # nn.AvgPool2d() is built without a kernel size -- the reference raises this TypeError for every
            # UNetModel(conv_resample=False) with more than one resolution level, so there is no behaviour to match
            raise TypeError("AvgPool2d.__init__() missing 1 required positional argument: 'kernel_size' "
                            "(conv_resample=False cannot be constructed in the reference either: unet_openai.py:209)")
        self.op = conv_nd(dims, channels, channels, 3, stride=2, padding=1)


class ResBlock(_Container):  # unet_openai.py:216-305
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False):
        super().__init__()
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, use_scale_shift_norm
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        # use_scale_shift_norm (unet_openai.py:257-260, 296-300; no shipped command sets it): the projection yields
        # [scale | shift], applied to the second GroupNorm's output instead of being added to conv #1's output
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm
                                                          else self.out_channels))
        self.cond_channels = self.emb_layers[1].out_features
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)


class AttentionBlock(_Container):  # unet_openai.py:308-342
    def __init__(self, channels, num_heads=1, use_checkpoint=False):
        super().__init__()
        self.channels, self.num_heads, self.use_checkpoint = channels, num_heads, use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttention()
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))


class QKVAttention(_Container):  # unet_openai.py:345-358 (no parameters)
    pass


class _Act:
    """Channels-last activation: fp32 tensor (GroupNorm / residual input) and its bf16 GEMM-operand copy."""
    __slots__ = ("c", "f32", "bf16")

    def __init__(self, c, f32=None, bf16=None):
        self.c, self.f32, self.bf16 = c, f32, bf16


class UNetModel(nn.Module):
    """unet_openai.py:361-575."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, z_dim=None, num_classes=None,
                 use_checkpoint=False, num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False, use_z=False):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("the latent score nets are 2-D")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_heads_upsample = num_heads_upsample

        time_embed_dim = model_channels * 4
        self.time_embed_dim = time_embed_dim
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), nn.SiLU(),
                                        linear(time_embed_dim, time_embed_dim))
        self.proj = None
        if use_z:
            self.proj = nn.Sequential(linear(z_dim, time_embed_dim), nn.SiLU(), linear(time_embed_dim, time_embed_dim))
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3,
                                                                           padding=1))])
        input_block_chans = [model_channels]
        ch = model_channels
        ds = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, dims=dims,
                                   use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm),
            AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads),
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm))
        self.output_blocks = nn.ModuleList([])
        self._cat_plan = []  # per output block: (channels of h, channels of the popped skip)
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                skip_ch = input_block_chans.pop()
                self._cat_plan.append((ch, skip_ch))
                layers = [ResBlock(ch + skip_ch, time_embed_dim, dropout, out_channels=model_channels * mult,
                                   dims=dims, use_checkpoint=use_checkpoint,
                                   use_scale_shift_norm=use_scale_shift_norm)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))
        self._packed: dict = {}
        self._res_blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
        self._res_index = {id(blk): i for i, blk in enumerate(self._res_blocks)}
        self._dropout_seed = None
        self._dropout_ctr = None

    @property
    def inner_dtype(self):  # unet_openai.py:531-536
        return next(self.input_blocks.parameters()).dtype

    # ------------------------------------------------------------------ packed-weight cache (keyed by version)
    def _cached(self, key, params, build):
        sig = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._packed[key] = (sig, val)
        return val

    def _w_conv(self, conv):
        return self._cached(id(conv), (conv.weight,), lambda: ops.pack_conv2d_weight(conv.weight))

    def _w_conv1d(self, conv):
        w = conv.weight
        return self._cached(id(conv), (w,), lambda: ops.pack_linear_weight(w.detach().reshape(w.shape[0], w.shape[1])))

    def _w_linear(self, lin):
        return self._cached(id(lin), (lin.weight,), lambda: ops.pack_linear_weight(lin.weight))

    def _w_emb2(self, with_z):
        """Second layer of the time MLP, optionally concatenated along K with the second layer of the z projection:
        emb = [SiLU(t1) | SiLU(z1)] @ [W_t2 | W_z2]^T + (b_t2 + b_z2)   (unet_openai.py:551-559)."""
        lt = self.time_embed[2]
        if not with_z:
            return self._cached("emb2", (lt.weight, lt.bias),
                                lambda: (ops.pack_linear_weight(lt.weight), lt.bias.detach().float().contiguous()))
        lz = self.proj[2]

        def build():
            w = torch.cat([lt.weight.detach(), lz.weight.detach()], dim=1).contiguous()
            return ops.pack_linear_weight(w), (lt.bias.detach() + lz.bias.detach()).float().contiguous()

        return self._cached("emb2z", (lt.weight, lt.bias, lz.weight, lz.bias), build)

    def _w_cond(self):
        """All per-block `emb_layers` linears (unet_openai.py:255-261) as ONE [sum(C_out), time_embed_dim] operand."""
        blocks = self._res_blocks
        params = tuple(b.emb_layers[1].weight for b in blocks) + tuple(b.emb_layers[1].bias for b in blocks)

        def build():
            total = sum(b.cond_channels for b in blocks)
            wpk = torch.empty((1, total, pad8(self.time_embed_dim)), dtype=torch.bfloat16, device=params[0].device)
            off, offsets = 0, {}
            for b in blocks:
                ops.pack_linear_weight(b.emb_layers[1].weight, out=wpk[:, off:off + b.cond_channels])
                offsets[id(b)] = off
                off += b.cond_channels
            bias = torch.cat([b.emb_layers[1].bias.detach().float() for b in blocks]).contiguous()
            return wpk, bias, offsets, total

        return self._cached("cond", params, build)

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _gn32_silu(x_f32, c, gn, act, mod_scale=None, mod_shift=None):
        b, h, w, _ = x_f32.shape
        st = torch.zeros((b, gn.num_groups, 2), dtype=torch.float64, device=x_f32.device)
        ops.group_stats(x_f32, c, gn.num_groups, st)
        a = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=x_f32.device)
        ops.groupnorm_apply(x_f32, c, st, gn.weight, gn.bias, groups=gn.num_groups, act=act, out=a, eps=gn.eps,
                            mod_scale=mod_scale, mod_shift=mod_shift)
        return a

    def _res_block(self, blk: ResBlock, x: _Act, cond, off, dst) -> _Act:
        """unet_openai.py:291-305 (eval mode: dropout is the identity)."""
        c_in, c_out = blk.channels, blk.out_channels
        a = self._gn32_silu(x.f32, c_in, blk.in_layers[0], L.ACT_SILU)
        conv1 = blk.in_layers[2]
        if blk.use_scale_shift_norm:   # out_norm(h) * (1 + scale) + shift, then SiLU (unet_openai.py:296-300)
            h = ops.conv_igemm(a, self._w_conv(conv1), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_out,
                               bias=conv1.bias)
            a2 = self._gn32_silu(h, c_out, blk.out_layers[0], L.ACT_SILU, mod_scale=cond[:, 0, 0, off:off + c_out],
                                 mod_shift=cond[:, 0, 0, off + c_out:off + 2 * c_out])
        else:
            h = ops.conv_igemm(a, self._w_conv(conv1), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_out,
                               bias=conv1.bias, rowbias=cond[:, :, :, off:off + c_out])
            a2 = self._gn32_silu(h, c_out, blk.out_layers[0], L.ACT_SILU)
        sk = blk.skip_connection
        if isinstance(sk, nn.Conv2d):
            k = sk.kernel_size[0]
            res = ops.conv_igemm(x.bf16, self._w_conv(sk), kind=L.CONV_S1, kh=k, kw=k, cin=c_in, cout=c_out,
                                 bias=sk.bias)
        else:
            res = x.f32
        conv2 = blk.out_layers[3]
        of, ob = dst(c_out, a2.shape)
        ops.conv_igemm(a2, self._w_conv(conv2), kind=L.CONV_S1, kh=3, kw=3, cin=c_out, cout=c_out, bias=conv2.bias,
                       residual=res, out=of, out2=ob)
        return _Act(c_out, f32=of, bf16=ob)

    def _attention(self, blk: AttentionBlock, x: _Act, dst) -> _Act:
        """unet_openai.py:329-358: GroupNorm32 -> qkv (1x1) -> per-head softmax(q k^T / sqrt(ch)) v -> proj_out + x."""
        c = blk.channels
        a = self._gn32_silu(x.f32, c, blk.norm, L.ACT_NONE)
        qkv = ops.conv_igemm(a, self._w_conv1d(blk.qkv), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=3 * c,
                             bias=blk.qkv.bias)
        dh = c // blk.num_heads
        o = ops.softmax_attn(qkv, blk.num_heads, dh, 0, dh, 2 * dh, 3 * dh, 1.0 / math.sqrt(dh))
        of, ob = dst(c, a.shape)
        ops.conv_igemm(o, self._w_conv1d(blk.proj_out), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=c,
                       bias=blk.proj_out.bias, residual=x.f32, out=of, out2=ob)
        return _Act(c, f32=of, bf16=ob)

    # ------------------------------------------------------------------ forward
    def forward(self, x, timesteps, z=None, y=None, **kwargs):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise L.SbmError("UNetModel.forward needs CUDA tensors: the B200 path has no CPU fallback")
        # train() mode with dropout > 0 always takes the training plan (it owns the mask kernel), also under no_grad
        if (self.training and self.dropout > 0) or (torch.is_grad_enabled()
                                                    and any(p.requires_grad for p in self.parameters())):
            from .autograd_openai import unet_openai_forward_train
            return unet_openai_forward_train(self, x, timesteps, z, y)
        with torch.no_grad():
            # large batches run in slices (exact: GroupNorm32 and the attention are per-sample)
            b = x.shape[0]
            widest = x.shape[-2] * x.shape[-1] * self.model_channels * max(self.channel_mult) * 2
            mb = max(1, self.max_chunk_elems // widest)
            if b <= mb:
                return self._forward_infer(x, timesteps, z, y)
            out = torch.empty((b, self.out_channels, x.shape[-2], x.shape[-1]), dtype=torch.float32, device=x.device)
            for lo in range(0, b, mb):
                out[lo:lo + mb] = self._forward_infer(x[lo:lo + mb], timesteps[lo:lo + mb],
                                                      None if z is None else z[lo:lo + mb],
                                                      None if y is None else y[lo:lo + mb])
            return out

    max_chunk_elems = 1 << 29

    # ------------------------------------------------------------------ dropout stream
    _DRAWS_PER_FORWARD = 4096  # > number of ResBlocks: draw id = forward index * 4096 + ResBlock index

    def set_dropout_seed(self, seed: int) -> None:
        """Philox key of the dropout masks (default: torch.initial_seed() at the first training forward; give every
        data-parallel rank its own)."""
        self._dropout_seed = int(seed)

    def _dropout_draw(self):
        """-> (seed, device scalar holding this forward's draw base).  The counter is snapshotted and advanced on the
        stream (inside a captured graph too), so the backward of THIS forward regenerates the same masks."""
        dev = next(self.parameters()).device
        if getattr(self, "_dropout_seed", None) is None:
            self._dropout_seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + 0x5D) & 0xFFFFFFFFFFFFFFFF
        ctr = getattr(self, "_dropout_ctr", None)
        if ctr is None or ctr.device != dev:
            ctr = self._dropout_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        snap = ctr.clone()
        L.check(L.lib().sbm_train_tick(None, L.ptr(ctr), C.c_uint64(self._DRAWS_PER_FORWARD), L.stream_ptr()),
                "sbm_train_tick")
        return self._dropout_seed, snap

    def _label_rows(self, y, b):
        """`emb + self.label_emb(y)` (unet_openai.py:561-564): the looked-up rows enter the second time-MLP GEMM as a
        per-sample row bias (added before its SiLU).  The row gather itself is a torch index op."""
        if y.shape != (b,):
            raise ValueError(f"y must have shape ({b},), got {tuple(y.shape)}")
        return self.label_emb.weight.detach().float()[y.long()].contiguous().view(b, 1, 1, self.time_embed_dim)

    def _forward_infer(self, x, timesteps, z, y=None):
        b, m, hh, ww = x.shape
        if m != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {m}")
        dev = x.device
        x = x.contiguous().float()
        ted, mc = self.time_embed_dim, self.model_channels

        # --- embedding path: SiLU(emb) feeds every block, emb itself is never needed
        te = ops.time_embed(timesteps.contiguous().float(), mc, 1)
        with_z = z is not None
        if with_z:
            assert self.proj is not None
        hcat = torch.empty((b, 1, 1, (2 if with_z else 1) * ted), dtype=torch.bfloat16, device=dev)
        l0 = self.time_embed[0]
        ops.conv_igemm(te, self._w_linear(l0), kind=L.CONV_S1, kh=1, kw=1, cin=mc, cout=ted, bias=l0.bias,
                       act=L.ACT_SILU, out=hcat[..., :ted])
        if with_z:
            zb, _ = ops.nchw_to_nhwc(z.contiguous().float().view(b, -1, 1, 1))
            p0 = self.proj[0]
            ops.conv_igemm(zb, self._w_linear(p0), kind=L.CONV_S1, kh=1, kw=1, cin=z.shape[1], cout=ted, bias=p0.bias,
                           act=L.ACT_SILU, out=hcat[..., ted:])
        w2, b2 = self._w_emb2(with_z)
        emb_act = ops.conv_igemm(hcat, w2, kind=L.CONV_S1, kh=1, kw=1, cin=hcat.shape[-1], cout=ted, bias=b2,
                                 act=L.ACT_SILU, out_dtype=torch.bfloat16,
                                 rowbias=self._label_rows(y, b) if y is not None else None)
        wc, bc, offs, total = self._w_cond()
        cond = ops.conv_igemm(emb_act, wc, kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=total, bias=bc)

        # --- concat buffers: output block j reads cat_j = [h (plan[j][0] channels) | skip (plan[j][1] channels)]
        n_out = len(self.output_blocks)
        cats = [None] * n_out

        def cat_for(j, shape_bhw):
            if cats[j] is None:
                ch_h, ch_s = self._cat_plan[j]
                bb, h_, w_ = shape_bhw
                cats[j] = (torch.empty((bb, h_, w_, ch_h + ch_s), dtype=torch.float32, device=dev),
                           torch.empty((bb, h_, w_, ch_h + ch_s), dtype=torch.bfloat16, device=dev))
            return cats[j]

        def skip_dst(i):  # i-th pushed feature map -> second channel range of its consumer's concat buffer
            j = n_out - 1 - i

            def dst(c, shape):
                cf, cb = cat_for(j, shape[:3])
                ch_h = self._cat_plan[j][0]
                return cf[..., ch_h:ch_h + c], cb[..., ch_h:ch_h + c]
            return dst

        def h_dst(j):  # input h of output block j -> first channel range of its concat buffer
            def dst(c, shape):
                cf, cb = cat_for(j, shape[:3])
                return cf[..., :c], cb[..., :c]
            return dst

        def plain_dst(c, shape):
            return (torch.empty((*shape[:3], pad8(c)), dtype=torch.float32, device=dev),
                    torch.empty((*shape[:3], pad8(c)), dtype=torch.bfloat16, device=dev))

        # --- input blocks
        xb, _ = ops.nchw_to_nhwc(x)
        stem = self.input_blocks[0][0]
        of, ob = skip_dst(0)(mc, (b, hh, ww))
        ops.conv_igemm(xb, self._w_conv(stem), kind=L.CONV_S1, kh=3, kw=3, cin=m, cout=mc, bias=stem.bias, out=of,
                       out2=ob)
        cur = _Act(mc, f32=of, bf16=ob)
        for i, block in enumerate(self.input_blocks):
            if i == 0:
                continue
            layers = list(block)
            for li, layer in enumerate(layers):
                dst = skip_dst(i) if li == len(layers) - 1 else plain_dst
                if isinstance(layer, ResBlock):
                    cur = self._res_block(layer, cur, cond, offs[id(layer)], dst)
                elif isinstance(layer, AttentionBlock):
                    cur = self._attention(layer, cur, dst)
                else:  # Downsample: 3x3 stride-2 conv (unet_openai.py:207)
                    h_, w_ = cur.bf16.shape[1:3]
                    of, ob = dst(cur.c, (b, h_ // 2, w_ // 2))
                    ops.conv_igemm(cur.bf16, self._w_conv(layer.op), kind=L.CONV_S2, kh=3, kw=3, cin=cur.c, cout=cur.c,
                                   bias=layer.op.bias, out=of, out2=ob)
                    cur = _Act(cur.c, f32=of, bf16=ob)

        # --- middle block (its output is the `h` half of the first concat buffer)
        mb = list(self.middle_block)
        cur = self._res_block(mb[0], cur, cond, offs[id(mb[0])], plain_dst)
        cur = self._attention(mb[1], cur, plain_dst)
        cur = self._res_block(mb[2], cur, cond, offs[id(mb[2])], h_dst(0))

        # --- output blocks
        for j, block in enumerate(self.output_blocks):
            cf, cb = cats[j]
            cur = _Act(cf.shape[-1], f32=cf, bf16=cb)
            layers = list(block)
            for li, layer in enumerate(layers):
                last = li == len(layers) - 1
                dst = (h_dst(j + 1) if j + 1 < n_out else plain_dst) if last else plain_dst
                if isinstance(layer, ResBlock):
                    cur = self._res_block(layer, cur, cond, offs[id(layer)], dst)
                elif isinstance(layer, AttentionBlock):
                    cur = self._attention(layer, cur, dst)
                else:  # Upsample: nearest 2x + 3x3 conv (unet_openai.py:185-187; a UNetModel with conv_resample=False
                    # never gets here: its Downsample cannot be constructed, see above)
                    up = ops.upsample_nearest2x(cur.bf16, cur.c)
                    of, ob = dst(cur.c, up.shape)
                    ops.conv_igemm(up, self._w_conv(layer.conv), kind=L.CONV_S1, kh=3, kw=3, cin=cur.c, cout=cur.c,
                                   bias=layer.conv.bias, out=of, out2=ob)
                    cur = _Act(cur.c, f32=of, bf16=ob)

        # --- out: GroupNorm32 -> SiLU -> conv3x3 -> NCHW fp32 (unet_openai.py:525-529)
        a = self._gn32_silu(cur.f32, cur.c, self.out[0], L.ACT_SILU)
        oc = self.out[2]
        return ops.conv_igemm(a, self._w_conv(oc), kind=L.CONV_S1, kh=3, kw=3, cin=cur.c, cout=self.out_channels,
                              bias=oc.bias, nchw=True)


class SuperResModel(UNetModel):  # unet_openai.py:578-592
    def __init__(self, in_channels, *args, **kwargs):
        super().__init__(in_channels * 2, *args, **kwargs)

    def forward(self, x, timesteps, low_res=None, **kwargs):
        """unet_openai.py:587-593: the low-resolution conditioning image is up-sampled (nearest) to x's extent and
        concatenated along the channels; the net itself is UNetModel's kernel plan (2 * in_channels inputs).  No shipped
        command of the reference instantiates this class; the input glue is two torch tensor ops."""
        if low_res is not None:
            up = torch.nn.functional.interpolate(low_res, tuple(x.shape[-2:]), mode="nearest")
            x = torch.cat([x, up], dim=1)
        return super().forward(x, timesteps, **kwargs)
