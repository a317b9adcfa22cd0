"""Training-mode execution of `Unet` (forward that keeps what the backward needs + hand-written backward).

`loss.backward()` in the reference (train_lat_celebhq_unet_cont2.py:98-100) differentiates ~350 ATen ops; here the
whole score net is ONE autograd node: `UnetFn.forward` runs the same kernels as inference and records a tape of
backward closures, `UnetFn.backward` replays it.  Every gradient is computed by kernels of libsbmae_b200:
  * data gradients of convolutions  = the forward implicit-GEMM kernel on re-packed (transposed / flipped) weights
    (stride-2 conv <-> transposed conv swap roles);
  * weight gradients                = `sbm_conv_wgrad` (tcgen05, MN-major operands straight from channels-last tensors);
  * GroupNorm / GELU / depthwise / attention-core backward = csrc/backward.cu.
Parameter gradients are returned to autograd as fp32 tensors, so torch optimizers, DDP hooks and `FusedAdam` all work.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from .ops import pad8


class _Node:
    """Channels-last activation with its gradient accumulator."""
    __slots__ = ("c", "f32", "bf16", "stats", "g")

    def __init__(self, c, f32=None, bf16=None, stats=None):
        self.c, self.f32, self.bf16, self.stats, self.g = c, f32, bf16, stats, None


def _acc(node: _Node, g: torch.Tensor) -> None:
    if node.g is None:
        node.g = g
    else:
        ops.add(node.g, g, node.c, out=node.g)


class _Plan:
    def __init__(self, model):
        self.m = model
        self.tape = []
        self.pg = {}  # parameter -> gradient

    # ------------------------------------------------------------------ weight packs for the data gradients
    def _dg_conv_s1(self, conv):
        w = conv.weight
        o, i, kh, kw = w.shape
        kk = kh * kw

        def build():
            wd = w.detach().contiguous()
            flat = wd.view(-1)[kk - 1:]  # base at the LAST tap; s_tap = -1 walks the taps flipped
            return ops.pack_weight(flat, kk, i, o, -1, kk, i * kk)

        return self.m._cached((id(conv), "dg"), (w,), build)

    def _dg_as_convT(self, conv):  # data gradient of a stride-2 nn.Conv2d = transposed conv with the same tensor
        return self.m._cached((id(conv), "dg"), (conv.weight,), lambda: ops.pack_convT2d_weight(conv.weight))

    def _dg_as_conv(self, convT):  # data gradient of nn.ConvTranspose2d = stride-2 conv with the same tensor
        return self.m._cached((id(convT), "dg"), (convT.weight,), lambda: ops.pack_conv2d_weight(convT.weight))

    def _dg_linear(self, lin):
        w = lin.weight
        o, i = w.shape
        return self.m._cached((id(lin), "dg"), (w,), lambda: ops.pack_weight(w.detach().contiguous(), 1, i, o, 0, 1, i))

    def _slot(self, p):
        """Data-parallel training: the parameter's slot in the flat bucket buffer, so that the producing kernel writes
        the gradient there directly (no copy in `grad_ready`); None otherwise."""
        sink = getattr(self.m, "_grad_sink", None)
        return sink.grad_view(p) if sink is not None and p not in self.pg else None

    def _grad(self, p, g):
        sink = getattr(self.m, "_grad_sink", None)
        if sink is not None:  # data-parallel training: copy into the flat bucket buffer, all-reduce when a bucket fills
            # deferred weight-gradient unpacks must land (a) before their bucket goes out and (b) before grad_ready
            # COPIES a gradient that was not produced in its slot: a deferred unpack into a private tensor is still
            # unwritten at this point (round 2: the time-MLP / stem / last-conv weights went out as uninitialised memory,
            # tools/check_multi_gpu.py: rank-averaged gradients 8e-2 off the single-GPU ones)
            if sink.completes_bucket(p) or g.data_ptr() != sink.grad_view(p).data_ptr():
                ops.unpack_flush()
            g = sink.grad_ready(p, g)
        if p in self.pg:
            ops.unpack_flush()   # both contributions must be materialised before they are summed
            self.pg[p] = self.pg[p] + g
        else:
            self.pg[p] = g

    # ------------------------------------------------------------------ generic conv backward pieces
    def _conv_s1_bwd(self, conv, x_b, dy_b, dy_f32_for_bias, cin, cout, k, want_dx=True, bias_grad=None):
        """gradients of y = conv_kxk_same(x) + bias.  Returns dx fp32 (or None)."""
        dwpk = ops.conv_wgrad(x_b, dy_b, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout)
        self._grad(conv.weight, ops.unpack_conv2d_wgrad(dwpk, conv.weight, self._slot(conv.weight)))
        if conv.bias is not None:
            self._grad(conv.bias, bias_grad if bias_grad is not None else ops.colsum(dy_f32_for_bias, cout))
        if not want_dx:
            return None
        return ops.conv_igemm(dy_b, self._dg_conv_s1(conv), kind=L.CONV_S1, kh=k, kw=k, cin=cout, cout=cin)

    # ------------------------------------------------------------------ ConvNeXt block
    def convnext(self, blk, xin: _Node, cond, cond_off, ldc, dcond, *, want_f32=True, want_bf16=False,
                 want_stats=False, out_f32=None, out_bf16=None) -> _Node:
        m = self.m
        xf = xin.f32
        b, h, w, _ = xf.shape
        dev = xf.device
        c_in, c_hid, c_out = blk.dim, blk.hidden, blk.dim_out
        has_t = cond is not None and blk.mlp is not None
        st1 = m._stats(b, dev)
        hdw = ops.dwconv7(xf, c_in, blk.ds_conv.weight, blk.ds_conv.bias, cond[:, :, :, cond_off:] if has_t else None,
                          ldc, st1)
        a1 = torch.empty((b, h, w, pad8(c_in)), dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(hdw, c_in, st1, blk.net[0].weight, blk.net[0].bias, out=a1)
        st2 = m._stats(b, dev)
        h2pre = torch.empty((b, h, w, pad8(c_hid)), dtype=torch.bfloat16, device=dev)
        h2 = ops.conv_igemm(a1, m._w_conv(blk.net[1]), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_hid,
                            bias=blk.net[1].bias, act=L.ACT_GELU, out_dtype=torch.bfloat16, stats=st2, out2=h2pre,
                            out2_preact=True)
        a2 = torch.empty((b, h, w, pad8(c_hid)), dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(h2, c_hid, st2, blk.net[3].weight, blk.net[3].bias, out=a2)
        del h2
        has_res = isinstance(blk.res_conv, nn.Conv2d)
        if has_res:
            res = ops.conv_igemm(xin.bf16, m._w_conv(blk.res_conv), kind=L.CONV_S1, kh=1, kw=1, cin=c_in, cout=c_out,
                                 bias=blk.res_conv.bias)
        else:
            res = xf
        st_out = m._stats(b, dev) if want_stats else None
        of = ob = None
        if want_f32:
            of = out_f32 if out_f32 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.float32, device=dev)
            if want_bf16:
                ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.bfloat16,
                                                                        device=dev)
            ops.conv_igemm(a2, m._w_conv(blk.net[4]), kind=L.CONV_S1, kh=3, kw=3, cin=c_hid, cout=c_out,
                           bias=blk.net[4].bias, residual=res, out=of, stats=st_out, out2=ob)
        else:
            ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.bfloat16,
                                                                    device=dev)
            ops.conv_igemm(a2, m._w_conv(blk.net[4]), kind=L.CONV_S1, kh=3, kw=3, cin=c_hid, cout=c_out,
                           bias=blk.net[4].bias, residual=res, out=ob, stats=st_out)
        out = _Node(c_out, f32=of, bf16=ob, stats=st_out)
        xin_b = xin.bf16

        def bwd():
            g = out.g
            _, g_b = ops.add(g, None, c_out, want_bf16=True)
            db_out = ops.colsum(g, c_out)
            da2 = self._conv_s1_bwd(blk.net[4], a2, g_b, None, c_hid, c_out, 3, bias_grad=db_out)
            _, dpre_b, dg2, dbeta2 = ops.groupnorm_bwd(h2pre, da2, c_hid, st2, blk.net[3].weight, in_act=L.ACT_GELU,
                                                       want_f32=False, want_bf16=True)
            self._grad(blk.net[3].weight, dg2)
            self._grad(blk.net[3].bias, dbeta2)
            da1 = self._conv_s1_bwd(blk.net[1], a1, dpre_b, dpre_b, c_in, c_hid, 3)
            dhdw, _, dg1, dbeta1 = ops.groupnorm_bwd(hdw, da1, c_in, st1, blk.net[0].weight)
            self._grad(blk.net[0].weight, dg1)
            self._grad(blk.net[0].bias, dbeta1)
            dw_dw, db_dw = ops.dwconv7_wgrad(xf, dhdw, c_in, dcond[:, :, :, cond_off:] if has_t else None, ldc)
            self._grad(blk.ds_conv.weight, dw_dw)
            self._grad(blk.ds_conv.bias, db_dw)
            if has_res:
                dres = self._conv_s1_bwd(blk.res_conv, xin_b, g_b, None, c_in, c_out, 1, bias_grad=db_out)
            else:
                dres = g
            dx = ops.dwconv7_bwd_input(dhdw, c_in, blk.ds_conv.weight, addend=dres)
            _acc(xin, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ ResnetBlock (unet_model.py:67-90)
    def resnet(self, blk, xin: _Node, cond, cond_off, ldc, dcond, *, want_f32=True, want_bf16=False,
               want_stats=False, out_f32=None, out_bf16=None) -> _Node:
        """Block 1 = conv3x3 -> GroupNorm(groups) -> SiLU, + time projection, Block 2 the same, + res_conv(x)
        (`Unet(use_convnext=False)`).  Same signature as `convnext`; the bf16 operand copy of the output is always
        written (the next block's convolution reads it)."""
        m = self.m
        xf = xin.f32
        b, h, w, _ = xf.shape
        dev = xf.device
        c_in, c_out, G = blk.dim, blk.dim_out, blk.groups
        has_t = cond is not None and blk.mlp is not None
        x_b = xin.bf16
        if x_b is None:
            _, x_b = ops.add(xf, None, c_in, want_bf16=True)
        p1, n1, p2, n2 = blk.block1.proj, blk.block1.norm, blk.block2.proj, blk.block2.norm
        h1 = ops.conv_igemm(x_b, m._w_conv(p1), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_out, bias=p1.bias)
        st1 = torch.zeros((b, G, 2), dtype=torch.float64, device=dev)
        ops.group_stats(h1, c_out, G, st1)
        a1 = torch.empty((b, h, w, pad8(c_out)), dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(h1, c_out, st1, n1.weight, n1.bias, groups=G, act=L.ACT_SILU, out=a1, eps=n1.eps,
                            post_add=cond[:, 0, 0, cond_off:cond_off + c_out] if has_t else None)
        h2 = ops.conv_igemm(a1, m._w_conv(p2), kind=L.CONV_S1, kh=3, kw=3, cin=c_out, cout=c_out, bias=p2.bias)
        st2 = torch.zeros((b, G, 2), dtype=torch.float64, device=dev)
        ops.group_stats(h2, c_out, G, st2)
        has_res = isinstance(blk.res_conv, nn.Conv2d)
        if has_res:
            res = ops.conv_igemm(x_b, m._w_conv(blk.res_conv), kind=L.CONV_S1, kh=1, kw=1, cin=c_in, cout=c_out,
                                 bias=blk.res_conv.bias)
        else:
            res = xf
        of = None
        if want_f32 or want_stats:
            of = out_f32 if out_f32 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.float32, device=dev)
        ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c_out)), dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(h2, c_out, st2, n2.weight, n2.bias, groups=G, act=L.ACT_SILU, residual=res, out=ob,
                            out_f32=of, eps=n2.eps)
        st_out = None
        if want_stats:   # statistics of the PreNorm GroupNorm(1, C) that follows (the ConvNeXt path gets them from the
            st_out = m._stats(b, dev)   # last convolution's epilogue; here the block ends in a norm, not a GEMM)
            ops.group_stats(of, c_out, 1, st_out)
        out = _Node(c_out, f32=of, bf16=ob, stats=st_out)

        def bwd():
            g = out.g
            dh2, dh2_b, dg2, db2 = ops.groupnorm_bwd(h2, g, c_out, st2, n2.weight, groups=G, want_f32=True,
                                                     want_bf16=True, eps=n2.eps, beta=n2.bias, out_act=L.ACT_SILU)
            self._grad(n2.weight, dg2)
            self._grad(n2.bias, db2)
            da1 = self._conv_s1_bwd(p2, a1, dh2_b, dh2, c_out, c_out, 3)
            if has_t:
                ops.colsum_per_sample(da1, c_out, dcond[:, 0, 0, cond_off:cond_off + c_out])
            dh1, dh1_b, dg1, db1 = ops.groupnorm_bwd(h1, da1, c_out, st1, n1.weight, groups=G, want_f32=True,
                                                     want_bf16=True, eps=n1.eps, beta=n1.bias, out_act=L.ACT_SILU)
            self._grad(n1.weight, dg1)
            self._grad(n1.bias, db1)
            dx = self._conv_s1_bwd(p1, x_b, dh1_b, dh1, c_in, c_out, 3)
            if has_res:
                _, g_b = ops.add(g, None, c_out, want_bf16=True)
                dres = self._conv_s1_bwd(blk.res_conv, x_b, g_b, g, c_in, c_out, 1)
            else:
                dres = g
            ops.add(dx, dres, c_in, out=dx)
            _acc(xin, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ Residual(PreNorm(LinearAttention))
    def linear_attention(self, mod, x: _Node, *, out_f32=None, out_bf16=None, want_bf16=True) -> _Node:
        m = self.m
        pre, att = mod.fn, mod.fn.fn
        xf = x.f32
        b, h, w, _ = xf.shape
        dev = xf.device
        c = x.c
        a = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(xf, c, x.stats, pre.norm.weight, pre.norm.bias, out=a)
        hid = att.heads * att.dim_head
        qkv = ops.conv_igemm(a, m._w_conv(att.to_qkv), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=3 * hid)
        o = ops.linear_attn(qkv, att.heads, att.scale)
        st = m._stats(b, dev)
        y = ops.conv_igemm(o, m._w_conv(att.to_out[0]), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=c,
                           bias=att.to_out[0].bias, stats=st)
        of = out_f32 if out_f32 is not None else torch.empty((b, h, w, pad8(c)), dtype=torch.float32, device=dev)
        ob = None
        if want_bf16:
            ob = out_bf16 if out_bf16 is not None else torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=dev)
        ops.groupnorm_apply(y, c, st, att.to_out[1].weight, att.to_out[1].bias, residual=xf, out=ob, out_f32=of)
        out = _Node(c, f32=of, bf16=ob)
        xstats = x.stats

        def bwd():
            g = out.g
            dy, dy_b, dgo, dbo = ops.groupnorm_bwd(y, g, c, st, att.to_out[1].weight, want_f32=True, want_bf16=True)
            self._grad(att.to_out[1].weight, dgo)
            self._grad(att.to_out[1].bias, dbo)
            do = self._conv_s1_bwd(att.to_out[0], o, dy_b, dy, hid, c, 1)
            dqkv_b = ops.linear_attn_bwd(qkv, do, att.heads, att.scale)
            da = self._conv_s1_bwd(att.to_qkv, a, dqkv_b, None, c, 3 * hid, 1)
            dx, _, dgn, dbn = ops.groupnorm_bwd(xf, da, c, xstats, pre.norm.weight, addend=g)
            self._grad(pre.norm.weight, dgn)
            self._grad(pre.norm.bias, dbn)
            _acc(x, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ Residual(PreNorm(Attention))
    def mid_attention(self, mod, x: _Node) -> _Node:
        m = self.m
        pre, att = mod.fn, mod.fn.fn
        xf = x.f32
        b, h, w, _ = xf.shape
        c = x.c
        a = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=xf.device)
        ops.groupnorm_apply(xf, c, x.stats, pre.norm.weight, pre.norm.bias, out=a)
        hid = att.heads * att.dim_head
        qkv = ops.conv_igemm(a, m._w_conv(att.to_qkv), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=3 * hid)
        o = ops.softmax_attn(qkv, att.heads, att.dim_head, 0, hid, 2 * hid, att.dim_head, att.scale)
        y = ops.conv_igemm(o, m._w_conv(att.to_out), kind=L.CONV_S1, kh=1, kw=1, cin=hid, cout=c,
                           bias=att.to_out.bias, residual=xf)
        out = _Node(c, f32=y)
        xstats = x.stats

        def bwd():
            g = out.g
            _, g_b = ops.add(g, None, c, want_bf16=True)
            do = self._conv_s1_bwd(att.to_out, o, g_b, g, hid, c, 1)
            dqkv_b = ops.softmax_attn_bwd(qkv, do, att.heads, att.dim_head, 0, hid, 2 * hid, att.dim_head, att.scale,
                                          3 * hid)
            da = self._conv_s1_bwd(att.to_qkv, a, dqkv_b, None, c, 3 * hid, 1)
            dx, _, dgn, dbn = ops.groupnorm_bwd(xf, da, c, xstats, pre.norm.weight, addend=g)
            self._grad(pre.norm.weight, dgn)
            self._grad(pre.norm.bias, dbn)
            _acc(x, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ whole network
    def forward(self, x, time):
        m = self.m
        b, mch, hh, ww = x.shape
        dev = x.device
        n_levels = len(m.downs)
        block = self.convnext if m.use_convnext else self.resnet
        m._arena = torch.zeros((3 * len(m._time_blocks) + 2 * n_levels + 12, b, 2), dtype=torch.float64, device=dev)
        m._arena_next = 0
        # ---- time path
        te = ops.time_embed(time, m.dim, 0)
        lin1, lin3 = m.time_mlp[1], m.time_mlp[3]
        t1pre = torch.empty((b, 1, 1, pad8(m.time_dim)), dtype=torch.bfloat16, device=dev)
        t1 = ops.conv_igemm(te, m._w_linear(lin1), kind=L.CONV_S1, kh=1, kw=1, cin=m.dim, cout=m.time_dim,
                            bias=lin1.bias, act=L.ACT_GELU, out_dtype=torch.bfloat16, out2=t1pre, out2_preact=True)
        t2pre = torch.empty((b, 1, 1, pad8(m.time_dim)), dtype=torch.bfloat16, device=dev)
        cond_act = m._cond_act   # GELU (ConvNextBlock.mlp) / SiLU (ResnetBlock.mlp) in front of the time projections
        tg = ops.conv_igemm(t1, m._w_linear(lin3), kind=L.CONV_S1, kh=1, kw=1, cin=m.time_dim, cout=m.time_dim,
                            bias=lin3.bias, act=cond_act, out_dtype=torch.bfloat16, out2=t2pre, out2_preact=True)
        wc, bc, offs, total = m._w_cond()
        cond = ops.conv_igemm(tg, wc, kind=L.CONV_S1, kh=1, kw=1, cin=m.time_dim, cout=total, bias=bc)
        ldc = cond.stride(2)
        dcond = torch.zeros_like(cond)

        def bwd_time():
            td = m.time_dim
            _, dc_b = ops.add(dcond, None, total, want_bf16=True)
            dbc = ops.colsum(dcond, total)
            dwc = ops.conv_wgrad(tg, dc_b, kind=L.CONV_S1, kh=1, kw=1, cin=td, cout=total)
            for blk in m._time_blocks:
                o = offs[id(blk)]
                lin = blk.mlp[1]
                self._grad(lin.weight, ops.unpack_linear_wgrad(dwc[:, o:o + blk.cond_channels], lin.weight))
                self._grad(lin.bias, dbc[o:o + blk.cond_channels].clone())

            # pack W_cat^T once per parameter version: [1][td][sum(C)]
            params = tuple(blk.mlp[1].weight for blk in m._time_blocks)

            def build_t():
                wcat = torch.cat([blk.mlp[1].weight.detach() for blk in m._time_blocks], dim=0).contiguous()  # [sumC, td]
                return ops.pack_weight(wcat, 1, td, total, 0, 1, td)

            wct = m._cached("cond_dg", params, build_t)
            dtg = ops.conv_igemm(dc_b, wct, kind=L.CONV_S1, kh=1, kw=1, cin=total, cout=td)
            d2f, d2b = ops.act_bwd(dtg, t2pre, td, cond_act, want_f32=True, want_bf16=True)
            dw3 = ops.conv_wgrad(t1, d2b, kind=L.CONV_S1, kh=1, kw=1, cin=td, cout=td)
            self._grad(lin3.weight, ops.unpack_linear_wgrad(dw3, lin3.weight, self._slot(lin3.weight)))
            self._grad(lin3.bias, ops.colsum(d2f, td))
            dt1 = ops.conv_igemm(d2b, self._dg_linear(lin3), kind=L.CONV_S1, kh=1, kw=1, cin=td, cout=td)
            d1f, d1b = ops.act_bwd(dt1, t1pre, td, L.ACT_GELU, want_f32=True, want_bf16=True)
            dw1 = ops.conv_wgrad(te, d1b, kind=L.CONV_S1, kh=1, kw=1, cin=m.dim, cout=td)
            self._grad(lin1.weight, ops.unpack_linear_wgrad(dw1, lin1.weight, self._slot(lin1.weight)))
            self._grad(lin1.bias, ops.colsum(d1f, td))

        self.tape.append(bwd_time)  # runs LAST in the backward (tape is replayed in reverse)

        # ---- stem
        a0 = ops.stem_im2col(x, 7, 7)
        c0 = m.init_dim
        need_b = isinstance(m.downs[0][0].res_conv, nn.Conv2d)
        x0b = torch.empty((b, hh, ww, pad8(c0)), dtype=torch.bfloat16, device=dev) if need_b else None
        x0 = ops.conv_igemm(a0, m._w_stem(), kind=L.CONV_S1, kh=1, kw=1, cin=mch * 49, cout=c0, bias=m.init_conv.bias,
                            out2=x0b)
        cur = _Node(c0, f32=x0, bf16=x0b)
        stem_node = cur

        def bwd_stem():
            g = stem_node.g
            _, g_b = ops.add(g, None, c0, want_bf16=True)
            dw = ops.conv_wgrad(a0, g_b, kind=L.CONV_S1, kh=1, kw=1, cin=mch * 49, cout=c0)
            wflat = m.init_conv.weight.detach().reshape(c0, -1)
            self._grad(m.init_conv.weight, ops.unpack_linear_wgrad(dw, wflat).view_as(m.init_conv.weight))
            self._grad(m.init_conv.bias, ops.colsum(g, c0))

        self.tape.append(bwd_stem)

        skips = []
        for lv, (block1, block2, attn, down) in enumerate(m.downs):
            cur = block(block1, cur, cond, offs[id(block1)], ldc, dcond)
            cur = block(block2, cur, cond, offs[id(block2)], ldc, dcond, want_stats=True)
            c = cur.c
            h, w = cur.f32.shape[1:3]
            if lv >= 1:
                cat_f = torch.empty((b, h, w, 2 * c), dtype=torch.float32, device=dev)
                cat_b = torch.empty((b, h, w, 2 * c), dtype=torch.bfloat16, device=dev)
                cur = self.linear_attention(attn, cur, out_f32=cat_f[..., c:], out_bf16=cat_b[..., c:])
                skips.append((cat_f, cat_b, c, cur))
            else:
                cur = self.linear_attention(attn, cur)
            if isinstance(down, nn.Conv2d):
                nxt_need_b = isinstance(m.downs[lv + 1][0].res_conv, nn.Conv2d)
                ob = torch.empty((b, h // 2, w // 2, pad8(c)), dtype=torch.bfloat16, device=dev) if nxt_need_b else None
                xin = cur
                y = ops.conv_igemm(xin.bf16, m._w_conv(down), kind=L.CONV_S2, kh=4, kw=4, cin=c, cout=c, bias=down.bias,
                                   out2=ob)
                cur = _Node(c, f32=y, bf16=ob)

                def bwd_down(xin=xin, yn=cur, down=down, c=c):
                    g = yn.g
                    _, g_b = ops.add(g, None, c, want_bf16=True)
                    dw = ops.conv_wgrad(xin.bf16, g_b, kind=L.CONV_S2, kh=4, kw=4, cin=c, cout=c)
                    self._grad(down.weight, ops.unpack_conv2d_wgrad(dw, down.weight, self._slot(down.weight)))
                    self._grad(down.bias, ops.colsum(g, c))
                    dx = ops.conv_igemm(g_b, self._dg_as_convT(down), kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=c, cout=c)
                    _acc(xin, dx)

                self.tape.append(bwd_down)

        cur = block(m.mid_block1, cur, cond, offs[id(m.mid_block1)], ldc, dcond, want_stats=True)
        cur = self.mid_attention(m.mid_attn, cur)
        cat_f, cat_b, c, skip_node = skips.pop()
        cur = block(m.mid_block2, cur, cond, offs[id(m.mid_block2)], ldc, dcond, want_bf16=True,
                            out_f32=cat_f[..., :c], out_bf16=cat_b[..., :c])
        first_half = cur
        for u, (block1, block2, attn, up) in enumerate(m.ups):
            cat_node = _Node(2 * c, f32=cat_f, bf16=cat_b)

            def bwd_cat(cat_node=cat_node, first=first_half, skip=skip_node, c=c):
                g = cat_node.g
                _acc(first, g[..., :c])
                _acc(skip, g[..., c:])

            self.tape.append(bwd_cat)
            cur = block(block1, cat_node, cond, offs[id(block1)], ldc, dcond)
            cur = block(block2, cur, cond, offs[id(block2)], ldc, dcond, want_stats=True)
            cur = self.linear_attention(attn, cur)
            cu = cur.c
            xin = cur
            if skips:
                cat_f, cat_b, c, skip_node = skips.pop()
                ops.conv_igemm(xin.bf16, m._w_convT(up), kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=cu, cout=cu, bias=up.bias,
                               out=cat_f[..., :c], out2=cat_b[..., :c])
                cur = _Node(cu, f32=cat_f[..., :c], bf16=cat_b[..., :c])
                first_half = cur
            else:
                y = ops.conv_igemm(xin.bf16, m._w_convT(up), kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=cu, cout=cu,
                                   bias=up.bias)
                cur = _Node(cu, f32=y)

            def bwd_up(xin=xin, yn=cur, up=up, cu=cu):
                g = yn.g
                _, g_b = ops.add(g, None, cu, want_bf16=True)
                dw = ops.conv_wgrad(xin.bf16, g_b, kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=cu, cout=cu)
                self._grad(up.weight, ops.unpack_convT2d_wgrad(dw, up.weight, self._slot(up.weight)))
                self._grad(up.bias, ops.colsum(g, cu))
                dx = ops.conv_igemm(g_b, self._dg_as_conv(up), kind=L.CONV_S2, kh=4, kw=4, cin=cu, cout=cu)
                _acc(xin, dx)

            self.tape.append(bwd_up)

        fin = m.final_conv[0]
        cur = block(fin, cur, None, 0, 0, None, want_f32=False)
        last = m.final_conv[1]
        xf_node = cur
        out = ops.conv_igemm(cur.bf16, m._w_conv(last), kind=L.CONV_S1, kh=1, kw=1, cin=fin.dim_out, cout=m.out_dim,
                             bias=last.bias, nchw=True)

        def bwd_last(dout):
            d_b, d_f = ops.nchw_to_nhwc(dout, want_f32=True)
            dw = ops.conv_wgrad(xf_node.bf16, d_b, kind=L.CONV_S1, kh=1, kw=1, cin=fin.dim_out, cout=m.out_dim)
            self._grad(last.weight, ops.unpack_conv2d_wgrad(dw, last.weight, self._slot(last.weight)))
            self._grad(last.bias, ops.colsum(d_f, m.out_dim))
            xf_node.g = ops.conv_igemm(d_b, self._dg_conv_s1(last), kind=L.CONV_S1, kh=1, kw=1, cin=m.out_dim,
                                       cout=fin.dim_out)

        self.bwd_last = bwd_last
        return out

    def backward(self, dout):
        sink = getattr(self.m, "_grad_sink", None)
        if sink is not None:
            sink.begin()
        # every zero-initialised accumulator of this pass comes out of ONE zeroed buffer (one fill launch)
        ops.arena_begin(getattr(self.m, "_zero_arena_words", 0), dout.device)
        ops.unpack_begin()   # the ~100 weight-gradient unpacks of the pass leave as one launch (per bucket)
        try:
            self.bwd_last(dout.contiguous().float())
            for fn in reversed(self.tape):
                fn()
        finally:
            ops.unpack_end()
            self.m._zero_arena_words = ops.arena_end()
        self.tape = []
        if sink is not None:
            sink.finish()


class UnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, time, *params):
        if x.requires_grad or time.requires_grad:
            # the hand-written backward produces PARAMETER gradients only; silently returning a zero input gradient
            # would corrupt e.g. a score Jacobian or an encoder trained through the latent
            raise L.SbmError("Unet: gradients w.r.t. the input latent / time are not implemented on the B200 path "
                             "(detach the input, or differentiate the parameters only)")
        plan = _Plan(model)
        with torch.no_grad():
            out = plan.forward(x.contiguous().float(), time.contiguous().float())
        ctx.plan = plan
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        if plan is None:
            raise L.SbmError("Unet: the backward tape was already consumed (a second backward through the same forward, "
                             "e.g. retain_graph=True, is not supported: run the forward again)")
        with torch.no_grad():
            plan.backward(dout)
        grads = tuple(plan.pg.get(p) for p in ctx.params)
        ctx.plan = None
        return (None, None, None) + grads


def unet_forward_train(model, x, time):
    model._refresh_packed()  # one launch re-packs every GEMM operand the last optimizer step invalidated
    params = tuple(p for p in model.parameters() if p.requires_grad)
    return UnetFn.apply(model, x, time, *params)
