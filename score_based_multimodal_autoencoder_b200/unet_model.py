"""Drop-in for the reference's `unet_model.py` score network (`Unet`, unet_model.py:189-323).

Same constructor signature, same parameter names / shapes (`state_dict()` of the reference loads
unchanged) and the same `forward(x[B,M,D,D], time[B]) -> [B,M,D,D]` contract.
The sub-modules below are parameter containers that mirror the reference's module tree; the arithmetic
is NOT done by torch.nn: `Unet.forward` walks the tree and issues hand-written sm_100a kernels through
the C ABI (ops.py -> libsbmae_b200.so): tcgen05 implicit-GEMM convolutions, fused depthwise-7x7 +
time-condition + GroupNorm statistics, GroupNorm-apply, linear/softmax attention cores.

Internal layout: channels-last activations, fp32 residual stream, bf16 GEMM operands, fp32 accumulate,
GroupNorm statistics in fp64.  There is no CPU / eager fallback: a CPU tensor raises.
"""
from __future__ import annotations

import math
import os
from functools import partial

import torch
from torch import nn

from . import _lib as L
from . import ops
from .ops import pad8


def exists(x):
    return x is not None


def default(val, d):
    return val if exists(val) else (d() if callable(d) else d)


class _Container(nn.Module):
    """Parameter holder: executed by Unet.forward's kernel plan, not callable on its own."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is executed by Unet.forward's fused CUDA plan")


class Residual(_Container):  # unet_model.py:21-27
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class PreNorm(_Container):  # unet_model.py:179-187
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.GroupNorm(1, dim)


def Upsample(dim):  # unet_model.py:29-30
    return nn.ConvTranspose2d(dim, dim, 4, 2, 1)


def Downsample(dim):  # unet_model.py:32-33
    return nn.Conv2d(dim, dim, 4, 2, 1)


class SinusoidalPositionEmbeddings(_Container):  # unet_model.py:35-47
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class Block(_Container):  # unet_model.py:49-65
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()


class ResnetBlock(_Container):  # unet_model.py:67-90 (`Unet(use_convnext=False)`; no shipped command builds it)
    def __init__(self, dim, dim_out, *, time_emb_dim=None, groups=8):
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out)) if exists(time_emb_dim) else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out, self.hidden, self.groups = dim, dim_out, dim_out, groups
        self.cond_channels = dim_out   # the time projection is added to Block 1's OUTPUT (unet_model.py:84-87)


class ConvNextBlock(_Container):  # unet_model.py:92-124
    def __init__(self, dim, dim_out, *, time_emb_dim=None, mult=2, norm=True):
        super().__init__()
        self.mlp = nn.Sequential(nn.GELU(), nn.Linear(time_emb_dim, dim)) if exists(time_emb_dim) else None
        self.ds_conv = nn.Conv2d(dim, dim, 7, padding=3, groups=dim)
        self.net = nn.Sequential(
            nn.GroupNorm(1, dim) if norm else nn.Identity(),
            nn.Conv2d(dim, dim_out * mult, 3, padding=1),
            nn.GELU(),
            nn.GroupNorm(1, dim_out * mult),
            nn.Conv2d(dim_out * mult, dim_out, 3, padding=1),
        )
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out, self.hidden = dim, dim_out, dim_out * mult
        self.cond_channels = dim   # the time projection is added to the depthwise output (unet_model.py:118-121)


class Attention(_Container):  # unet_model.py:126-149
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        hidden = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


class LinearAttention(_Container):  # unet_model.py:151-177
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        hidden = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), nn.GroupNorm(1, dim))


class _Act:
    """A channels-last activation: fp32 residual-stream tensor and/or its bf16 GEMM-operand copy."""
    __slots__ = ("f32", "bf16", "c", "stats")

    def __init__(self, c, f32=None, bf16=None, stats=None):
        self.c, self.f32, self.bf16, self.stats = c, f32, bf16, stats


class Unet(nn.Module):
    """unet_model.py:189-323.  `use_convnext=True` (every shipped command) runs the tuned inference plan below;
    `use_convnext=False` (ResnetBlock, unet_model.py:67-90) runs on the same kernels through the tape plan of
    autograd.py (`_Plan.resnet`), for inference and training alike."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, with_time_emb=True,
                 resnet_block_groups=8, use_convnext=True, convnext_mult=2):
        super().__init__()
        if not with_time_emb:
            raise NotImplementedError("the score net is always time-conditioned (forward(x, t))")
        self.channels = channels
        self.dim = dim
        init_dim = default(init_dim, dim // 3 * 2)
        self.init_dim = init_dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        self.dim_mults = dim_mults
        dims = [init_dim, *map(lambda m: dim * m, dim_mults)]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.use_convnext = bool(use_convnext)
        if use_convnext:
            block_klass = partial(ConvNextBlock, mult=convnext_mult)
        else:
            block_klass = partial(ResnetBlock, groups=resnet_block_groups)
        # activation in front of every block's time projection (unet_model.py:98 GELU / :73 SiLU): applied ONCE to the
        # time-MLP output, all projections are one GEMM
        self._cond_act = L.ACT_GELU if use_convnext else L.ACT_SILU
        time_dim = dim * 4
        self.time_dim = time_dim
        self.time_mlp = nn.Sequential(SinusoidalPositionEmbeddings(dim), nn.Linear(dim, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        num_resolutions = len(in_out)
        for ind, (dim_in, dim_out) in enumerate(in_out):
            is_last = ind >= (num_resolutions - 1)
            self.downs.append(nn.ModuleList([
                block_klass(dim_in, dim_out, time_emb_dim=time_dim),
                block_klass(dim_out, dim_out, time_emb_dim=time_dim),
                Residual(PreNorm(dim_out, LinearAttention(dim_out))),
                Downsample(dim_out) if not is_last else nn.Identity(),
            ]))
        mid_dim = dims[-1]
        self.mid_block1 = block_klass(mid_dim, mid_dim, time_emb_dim=time_dim)
        self.mid_attn = Residual(PreNorm(mid_dim, Attention(mid_dim)))
        self.mid_block2 = block_klass(mid_dim, mid_dim, time_emb_dim=time_dim)
        for ind, (dim_in, dim_out) in enumerate(reversed(in_out[1:])):
            # the reference's `is_last` (unet_model.py:257) is never true here: every level upsamples
            self.ups.append(nn.ModuleList([
                block_klass(dim_out * 2, dim_in, time_emb_dim=time_dim),
                block_klass(dim_in, dim_in, time_emb_dim=time_dim),
                Residual(PreNorm(dim_in, LinearAttention(dim_in))),
                Upsample(dim_in),
            ]))
        out_dim = default(out_dim, channels)
        self.out_dim = out_dim
        self.final_conv = nn.Sequential(block_klass(dim, dim), nn.Conv2d(dim, out_dim, 1))
        self._packed: dict = {}
        self._time_blocks = [m for m in self.modules() if isinstance(m, (ConvNextBlock, ResnetBlock))
                             and m.mlp is not None]
        # bf16 storage for the GELU'd hidden activation between the two 3x3 convolutions of a block
        self.hidden_dtype = torch.bfloat16
        # inference: fold the block's first GroupNorm into its 3x3 convolution (depthwise output kept in bf16)
        self.fold_input_norm = True
        # inference: the 16x16 / 8x8 linear-attention blocks pass q | k | v from the to_qkv GEMM to the attention kernel
        # as bf16 (soft-max arithmetic stays fp32)
        self.qkv_bf16 = os.environ.get("SBM_QKV_BF16", "1") != "0"

    # ------------------------------------------------------------------ packed-weight cache
    @staticmethod
    def _sig(params):
        return tuple((p.data_ptr(), p._version) if p is not None else None for p in params)

    def _cached(self, key, params, build):
        sig = self._sig(params)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._packed[key] = (sig, val, params)
        return val

    def _refresh_packed(self):
        """Bring every stale bf16 weight pack up to date IN PLACE with one multi-tensor launch (an optimizer step
        invalidates all ~180 of them; re-packing one by one is 180 tiny launches per training step).  Packs whose
        source is not a live view of the parameters (concatenated copies, folded GroupNorm tables) are left to the lazy
        per-use rebuild of `_cached`."""
        todo = []
        for key, (sig, val, params) in self._packed.items():
            cur = self._sig(params)
            if cur == sig:
                continue
            packs = [val] if isinstance(val, torch.Tensor) else list(getattr(val[0], "_subpacks", ())) if isinstance(val, tuple) else []
            ok = bool(packs)
            storages = {p.untyped_storage().data_ptr() for p in params if p is not None}
            for t in packs:
                spec = getattr(t, "_pack_spec", None)
                if spec is None or spec[0].untyped_storage().data_ptr() not in storages:
                    ok = False
            if ok:
                todo.append((key, cur, val, params, packs))
        if not todo:
            return
        with torch.no_grad():
            ops.pack_weights_multi([t for _, _, _, _, packs in todo for t in packs])
            for key, cur, val, params, _ in todo:
                if isinstance(val, tuple) and hasattr(val[0], "_bias_sources"):  # concatenated bias of the cond GEMM
                    torch.cat([b.detach().float() for b in val[0]._bias_sources], out=val[1])
                self._packed[key] = (cur, val, params)

    def _w_conv(self, conv: nn.Conv2d):
        return self._cached(id(conv), (conv.weight,), lambda: ops.pack_conv2d_weight(conv.weight))

    def _w_conv_gn(self, conv: nn.Conv2d, gn: nn.GroupNorm):
        """GroupNorm(1,C) -> conv folded: (bf16 weights carrying gamma, border-class tables with beta / bias)."""
        return self._cached((id(conv), "gn"), (conv.weight, conv.bias, gn.weight, gn.bias),
                            lambda: ops.fold_groupnorm_conv(conv.weight, conv.bias, gn.weight, gn.bias))

    def _w_convT(self, conv: nn.ConvTranspose2d):
        return self._cached(id(conv), (conv.weight,), lambda: ops.pack_convT2d_weight(conv.weight))

    def _w_linear(self, lin: nn.Linear):
        return self._cached(id(lin), (lin.weight,), lambda: ops.pack_linear_weight(lin.weight))

    def _w_stem(self):
        w = self.init_conv.weight
        return self._cached("stem", (w,), lambda: ops.pack_linear_weight(w.detach().reshape(w.shape[0], -1)))

    def _w_cond(self):
        """All per-block time projections (unet_model.py:97-101) as ONE [sum(C), time_dim] GEMM operand."""
        blocks = self._time_blocks
        params = tuple(b.mlp[1].weight for b in blocks) + tuple(b.mlp[1].bias for b in blocks)

        def build():
            total = sum(b.cond_channels for b in blocks)
            wpk = torch.empty((1, total, pad8(self.time_dim)), dtype=torch.bfloat16, device=params[0].device)
            off = 0
            offsets = {}
            subpacks = []
            for b in blocks:
                subpacks.append(ops.pack_linear_weight(b.mlp[1].weight, out=wpk[:, off:off + b.cond_channels]))
                offsets[id(b)] = off
                off += b.cond_channels
            bias = torch.cat([b.mlp[1].bias.detach().float() for b in blocks]).contiguous()
            wpk._subpacks = subpacks                      # for the in-place multi-tensor refresh
            wpk._bias_sources = [b.mlp[1].bias for b in blocks]
            return wpk, bias, offsets, total

        return self._cached("cond", params, build)

    # ------------------------------------------------------------------ building blocks
    def _stats(self, b, dev):
        """One (sum, sumsq) slot per GroupNorm site, carved from an arena zeroed ONCE per forward."""
        i = self._arena_next
        self._arena_next += 1
        if i >= self._arena.shape[0]:
            return torch.zeros((b, 2), dtype=torch.float64, device=dev)
        return self._arena[i]

    def _convnext(self, blk: ConvNextBlock, x: _Act, cond, cond_off, ldc, *, want_f32=True, want_bf16=False,
                  want_stats=False, out_f32=None, out_bf16=None) -> _Act:
        xf = x.f32
        b, h, w, _ = xf.shape
        dev = xf.device
        c_in, c_hid, c_out = blk.dim, blk.hidden, blk.dim_out
        st1 = self._stats(b, dev)
        cptr = cond[:, :, :, cond_off:] if (cond is not None and blk.mlp is not None) else None
        st2 = self._stats(b, dev)
        if self.fold_input_norm and isinstance(blk.net[0], nn.GroupNorm) and h == w and w <= 16:
            # depthwise output stored once as bf16 (statistics from the unrounded values); the first GroupNorm is
            # folded into the 3x3 convolution like the second one
            hdw = ops.dwconv7(xf, c_in, blk.ds_conv.weight, blk.ds_conv.bias, cptr, ldc, st1, out_dtype=torch.bfloat16)
            w1, tab1 = self._w_conv_gn(blk.net[1], blk.net[0])
            h2 = ops.conv_igemm(hdw, w1, kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_hid, act=L.ACT_GELU,
                                out_dtype=self.hidden_dtype, stats=st2, gn_stats=st1, gn_tab=tab1,
                                gn_eps=blk.net[0].eps)
        else:
            hdw = ops.dwconv7(xf, c_in, blk.ds_conv.weight, blk.ds_conv.bias, cptr, ldc, st1)
            a1 = torch.empty((b, h, w, pad8(c_in)), dtype=torch.bfloat16, device=dev)
            ops.groupnorm_apply(hdw, c_in, st1, blk.net[0].weight, blk.net[0].bias, out=a1)
            h2 = ops.conv_igemm(a1, self._w_conv(blk.net[1]), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_hid,
                                bias=blk.net[1].bias, act=L.ACT_GELU, out_dtype=self.hidden_dtype, stats=st2)
        # the second GroupNorm is folded into the last convolution: it consumes the raw GELU output (bf16) with
        # gamma-carrying weights and applies mean / rstd / beta in its epilogue (no GroupNorm-apply pass)
        w4, tab4 = self._w_conv_gn(blk.net[4], blk.net[3])
        gn_kw = dict(gn_stats=st2, gn_tab=tab4, gn_eps=blk.net[3].eps)
        if isinstance(blk.res_conv, nn.Conv2d):
            res = ops.conv_igemm(x.bf16, self._w_conv(blk.res_conv), kind=L.CONV_S1, kh=1, kw=1, cin=c_in, cout=c_out,
                                 bias=blk.res_conv.bias)
        else:
            res = xf
        st_out = self._stats(b, dev) if want_stats else None
        if want_f32:
            of = out_f32 if out_f32 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.float32,
                                                                  device=dev)
            ob = None
            if want_bf16:
                ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.bfloat16,
                                                                        device=dev)
            ops.conv_igemm(h2, w4, kind=L.CONV_S1, kh=3, kw=3, cin=c_hid, cout=c_out, residual=res, out=of,
                           stats=st_out, out2=ob, **gn_kw)
            return _Act(c_out, f32=of, bf16=ob, stats=st_out)
        ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.bfloat16,
                                                                device=dev)
        ops.conv_igemm(h2, w4, kind=L.CONV_S1, kh=3, kw=3, cin=c_hid, cout=c_out, residual=res, out=ob, stats=st_out,
                       **gn_kw)
        return _Act(c_out, bf16=ob, stats=st_out)

    def _qkv(self, pre: PreNorm, att, x: _Act, bf16_out: bool = False):
        """PreNorm GroupNorm -> to_qkv (1x1, no bias).  With the bf16 copy of the residual stream at hand the norm is folded
        into the GEMM (gamma in the weights, mean / rstd / beta in the epilogue); otherwise GroupNorm-apply + GEMM.
        bf16_out: qkv leaves the GEMM as bf16 (the 16x16 / 8x8 linear-attention kernel reads it; at those map sizes the
        GEMM and the attention kernel are bound by the HBM traffic of exactly this tensor)."""
        xf = x.f32
        b, h, w, _ = xf.shape
        c = x.c
        hid = att.heads * att.dim_head
        od = torch.bfloat16 if bf16_out else torch.float32
        if self.fold_input_norm and x.bf16 is not None:
            wq, tabq = self._w_conv_gn(att.to_qkv, pre.norm)
            return ops.conv_igemm(x.bf16, wq, kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=3 * hid, gn_stats=x.stats,
                                  gn_tab=tabq, gn_eps=pre.norm.eps, out_dtype=od)
        a = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=xf.device)
        ops.groupnorm_apply(xf, c, x.stats, pre.norm.weight, pre.norm.bias, out=a)
        return ops.conv_igemm(a, self._w_conv(att.to_qkv), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=3 * hid, out_dtype=od)

    def _linear_attention(self, mod: Residual, x: _Act, *, out_f32=None, out_bf16=None, want_bf16=True) -> _Act:
        pre: PreNorm = mod.fn
        att: LinearAttention = pre.fn
        xf = x.f32
        b, h, w, _ = xf.shape
        dev = xf.device
        c = x.c
        hid = att.heads * att.dim_head
        qkv = self._qkv(pre, att, x, bf16_out=self.qkv_bf16 and h * w in (64, 256) and att.dim_head == 32)
        o = ops.linear_attn(qkv, att.heads, att.scale)
        st = self._stats(b, dev)
        y = ops.conv_igemm(o, self._w_conv(att.to_out[0]), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=c,
                           bias=att.to_out[0].bias, stats=st)
        of = out_f32 if out_f32 is not None else torch.empty((b, h, w, pad8(c)), dtype=torch.float32, device=dev)
        ob = None
        if want_bf16:
            ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16,
                                                                    device=dev)
        ops.groupnorm_apply(y, c, st, att.to_out[1].weight, att.to_out[1].bias, residual=xf, out=ob, out_f32=of)
        return _Act(c, f32=of, bf16=ob)

    def _mid_attention(self, mod: Residual, x: _Act) -> _Act:
        pre: PreNorm = mod.fn
        att: Attention = pre.fn
        xf = x.f32
        c = x.c
        hid = att.heads * att.dim_head
        qkv = self._qkv(pre, att, x)
        o = ops.softmax_attn(qkv, att.heads, att.dim_head, 0, hid, 2 * hid, att.dim_head, att.scale)
        y = ops.conv_igemm(o, self._w_conv(att.to_out), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=c,
                           bias=att.to_out.bias, residual=xf)
        return _Act(c, f32=y)

    # ------------------------------------------------------------------ forward
    def forward(self, x, time=None):
        if not x.is_cuda:
            raise L.SbmError("Unet.forward needs CUDA tensors: the B200 path has no CPU fallback")
        # non-power-of-two extents are zero-padded symmetrically and the output cropped (unet_model.py:276-284, 318-322)
        pw = int((2 ** math.ceil(math.log2(x.shape[-1])) - x.shape[-1]) // 2)
        ph = int((2 ** math.ceil(math.log2(x.shape[-2])) - x.shape[-2]) // 2)
        if pw or ph:
            x = torch.nn.functional.pad(x, (pw, pw, ph, ph))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import unet_forward_train
            y = unet_forward_train(self, x, time)
        else:
            y = self._forward_chunked(x, time)
        if pw:
            y = y[..., pw:-pw]
        if ph:
            y = y[..., ph:-ph, :]
        return y

    # widest activation of one sample-chunk, in elements: batches beyond it (the 4k-64k sweep of BASELINE configs[4]) run
    # through the net in slices -- exact, since GroupNorm(1, C) and both attentions are per-sample
    max_chunk_elems = 1 << 29

    def max_batch(self, hh, ww):
        widest = hh * ww * max(self.init_dim, max(blk.hidden for blk in self._convnext_blocks()))
        return max(1, self.max_chunk_elems // widest)

    def _convnext_blocks(self):
        return [mod for mod in self.modules() if isinstance(mod, (ConvNextBlock, ResnetBlock))]

    @torch.no_grad()
    def _forward_chunked(self, x, time):
        b = x.shape[0]
        mb = self.max_batch(x.shape[-2], x.shape[-1])
        if b <= mb:
            return self._forward_infer(x, time)
        n = -(-b // mb)
        step = -(-b // n)
        step = -(-step // 256) * 256 if step >= 256 else step   # whole 256-sample blocks for the pixel-major tiling
        out = torch.empty((b, self.out_dim, x.shape[-2], x.shape[-1]), dtype=torch.float32, device=x.device)
        for lo in range(0, b, step):
            out[lo:lo + step] = self._forward_infer(x[lo:lo + step], time[lo:lo + step])
        return out

    @torch.no_grad()
    def _forward_infer(self, x, time):
        b, m, hh, ww = x.shape
        if m != self.channels:
            raise ValueError(f"expected {self.channels} latent channels, got {m}")
        for s in (hh, ww):
            if s & (s - 1):
                # forward() pads to the next power of two like the reference; an odd deficit (e.g. 7) stays unpadded
                # there too and crashes in its first down-sampling conv
                raise ValueError("latent extent must pad to a power of two (unet_model.py:276-284)")
        dev = x.device
        x = x.contiguous().float()
        time = time.contiguous().float()
        if not self.use_convnext:
            # ResnetBlock variant: the tape plan's forward (same kernels); the tape is dropped with the plan
            from .autograd import _Plan
            return _Plan(self).forward(x, time)
        n_levels = len(self.downs)
        self._arena = torch.zeros((3 * len(self._time_blocks) + 2 * n_levels + 12, b, 2), dtype=torch.float64,
                                  device=dev)
        self._arena_next = 0

        # --- time path: sinusoid -> Linear -> GELU -> Linear, then the per-block GELU -> Linear as ONE GEMM
        te = ops.time_embed(time, self.dim, 0)
        t1 = ops.conv_igemm(te, self._w_linear(self.time_mlp[1]), kind=L.CONV_S1, kh=1, kw=1, cin=self.dim,
                            cout=self.time_dim, bias=self.time_mlp[1].bias, act=L.ACT_GELU,
                            out_dtype=torch.bfloat16)
        tg = ops.conv_igemm(t1, self._w_linear(self.time_mlp[3]), kind=L.CONV_S1, kh=1, kw=1, cin=self.time_dim,
                            cout=self.time_dim, bias=self.time_mlp[3].bias, act=L.ACT_GELU,
                            out_dtype=torch.bfloat16)
        wc, bc, offs, total = self._w_cond()
        cond = ops.conv_igemm(tg, wc, kind=L.CONV_S1, kh=1, kw=1, cin=self.time_dim, cout=total, bias=bc)
        ldc = cond.stride(2)

        # --- stem 7x7 as im2col + GEMM
        a0 = ops.stem_im2col(x, 7, 7)
        c0 = self.init_dim
        need_b = isinstance(self.downs[0][0].res_conv, nn.Conv2d)
        x0b = torch.empty((b, hh, ww, pad8(c0)), dtype=torch.bfloat16, device=dev) if need_b else None
        x0 = ops.conv_igemm(a0, self._w_stem(), kind=L.CONV_S1, kh=1, kw=1, cin=m * 49, cout=c0,
                            bias=self.init_conv.bias, out2=x0b)
        cur = _Act(c0, f32=x0, bf16=x0b)

        skips = []
        for lv, (block1, block2, attn, down) in enumerate(self.downs):
            cur = self._convnext(block1, cur, cond, offs[id(block1)], ldc)
            cur = self._convnext(block2, cur, cond, offs[id(block2)], ldc, want_stats=True,
                                 want_bf16=self.fold_input_norm)
            c = cur.c
            h, w = cur.f32.shape[1:3]
            # the level's output is the skip: write it straight into the 2nd half of the up path's concat buffer
            # (level 0's skip is never consumed: the reference pops only n_levels-1 skips, unet_model.py:310-311)
            if lv >= 1:
                cat_f = torch.empty((b, h, w, 2 * c), dtype=torch.float32, device=dev)
                cat_b = torch.empty((b, h, w, 2 * c), dtype=torch.bfloat16, device=dev)
                cur = self._linear_attention(attn, cur, out_f32=cat_f[..., c:], out_bf16=cat_b[..., c:])
                skips.append((cat_f, cat_b, c))
            else:
                cur = self._linear_attention(attn, cur)
            if isinstance(down, nn.Conv2d):
                nxt_need_b = isinstance(self.downs[lv + 1][0].res_conv, nn.Conv2d)
                ob = torch.empty((b, h // 2, w // 2, pad8(c)), dtype=torch.bfloat16, device=dev) if nxt_need_b else None
                y = ops.conv_igemm(cur.bf16, self._w_conv(down), kind=L.CONV_S2, kh=4, kw=4, cin=c, cout=c,
                                   bias=down.bias, out2=ob)
                cur = _Act(c, f32=y, bf16=ob)

        cur = self._convnext(self.mid_block1, cur, cond, offs[id(self.mid_block1)], ldc, want_stats=True,
                             want_bf16=self.fold_input_norm)
        cur = self._mid_attention(self.mid_attn, cur)
        # mid_block2's output is the first half of the first concat buffer
        cat_f, cat_b, c = skips.pop()
        cur = self._convnext(self.mid_block2, cur, cond, offs[id(self.mid_block2)], ldc, want_bf16=True,
                             out_f32=cat_f[..., :c], out_bf16=cat_b[..., :c])
        for u, (block1, block2, attn, up) in enumerate(self.ups):
            cur = _Act(2 * c, f32=cat_f, bf16=cat_b)
            cur = self._convnext(block1, cur, cond, offs[id(block1)], ldc)
            cur = self._convnext(block2, cur, cond, offs[id(block2)], ldc, want_stats=True,
                                 want_bf16=self.fold_input_norm)
            cur = self._linear_attention(attn, cur)
            cu = cur.c
            h, w = cur.f32.shape[1:3]
            if skips:
                cat_f, cat_b, c = skips.pop()
                # fp32 half for the depthwise conv + bf16 operand copy for the next block's 1x1 res_conv
                ops.conv_igemm(cur.bf16, self._w_convT(up), kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=cu, cout=cu,
                               bias=up.bias, out=cat_f[..., :c], out2=cat_b[..., :c])
            else:
                y = ops.conv_igemm(cur.bf16, self._w_convT(up), kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=cu, cout=cu,
                                   bias=up.bias)
                cur = _Act(cu, f32=y)
        fin: ConvNextBlock = self.final_conv[0]
        cur = self._convnext(fin, cur, None, 0, 0, want_f32=False)
        out = ops.conv_igemm(cur.bf16, self._w_conv(self.final_conv[1]), kind=L.CONV_S1, kh=1, kw=1, cin=fin.dim_out,
                             cout=self.out_dim, bias=self.final_conv[1].bias, nchw=True)
        return out
