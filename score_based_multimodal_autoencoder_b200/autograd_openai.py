"""Training-mode execution of `UNetModel` (unet_openai.py:361-575): forward that keeps what the backward needs +
hand-written backward, as ONE autograd node — the z-conditioned DSM training of
train_lat_celebhq_unet_cont2_cond.py:95-124 (`loss_fn(..., z_cond=z)`).

Same structure as `autograd._Plan` (the ConvNeXt `Unet`): the forward issues the inference kernels and records a tape
of closures; the backward replays it.  Every gradient comes from libsbmae_b200 kernels:
  * data gradients of the 3x3 / 1x1 convolutions = the forward implicit GEMM on flipped / transposed weight packs;
    the 3x3 stride-2 down-sampling conv (unet_openai.py:207) back-propagates as a 3x3 transposed conv with
    output_padding 1 (four output-parity phases of the same kernel);
  * weight gradients = `sbm_conv_wgrad`; GroupNorm32 + SiLU backward = `sbm_groupnorm_bwd(out_act=SiLU)`;
  * the `h + emb_out[..., None, None]` row bias (unet_openai.py:303) back-propagates through a per-sample column sum
    into ONE gradient GEMM for all `emb_layers` projections; `QKVAttention` through `sbm_softmax_attn_bwd`; nearest-2x
    up-sampling through a 2x2 block sum.
Dropout (unet_openai.py:265, p = 0.1 in train_lat_celebhq_unet_cont2_cond.py:653) = `sbm_dropout` in place on the
GroupNorm32+SiLU output (the conv / weight-gradient operand) and, with the same Philox coordinates, on the gradient
coming back from the conv: the mask is regenerated, never stored.  The per-forward draw id lives in device memory
(`UNetModel._dropout_ctr`, snapshotted then advanced at the start of every training forward) so a captured CUDA graph
draws fresh masks on every replay.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib as L
from . import ops
from .autograd import _Node, _Plan, _acc
from .ops import pad8


class _OPlan(_Plan):
    # ------------------------------------------------------------------ small helpers
    def _gn32(self, x_f32, c, gn, act):
        b, h, w, _ = x_f32.shape
        st = torch.zeros((b, gn.num_groups, 2), dtype=torch.float64, device=x_f32.device)
        ops.group_stats(x_f32, c, gn.num_groups, st)
        a = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=x_f32.device)
        ops.groupnorm_apply(x_f32, c, st, gn.weight, gn.bias, groups=gn.num_groups, act=act, out=a, eps=gn.eps)
        return a, st

    def _gn32_bwd(self, gn, x_f32, dy, c, st, act, addend=None, want_bf16=False):
        dx, dx_b, dg, db = ops.groupnorm_bwd(x_f32, dy, c, st, gn.weight, groups=gn.num_groups, addend=addend,
                                             want_f32=True, want_bf16=want_bf16, eps=gn.eps, beta=gn.bias, out_act=act)
        self._grad(gn.weight, dg)
        self._grad(gn.bias, db)
        return dx, dx_b

    def _conv2d_bwd(self, conv, x_b, g_b, g_f32, cin, cout, k, kind=L.CONV_S1, want_dx=True):
        """gradients of a 2-D convolution (stride 1 'same' or 3x3 stride 2): weight, bias, and dx (fp32) if wanted."""
        dwpk = ops.conv_wgrad(x_b, g_b, kind=kind, kh=k, kw=k, cin=cin, cout=cout)
        self._grad(conv.weight, ops.unpack_conv2d_wgrad(dwpk, conv.weight, self._slot(conv.weight)))
        self._grad(conv.bias, ops.colsum(g_f32, cout))
        if not want_dx:
            return None
        if kind == L.CONV_S1:
            return ops.conv_igemm(g_b, self._dg_conv_s1(conv), kind=L.CONV_S1, kh=k, kw=k, cin=cout, cout=cin)
        return ops.conv_igemm(g_b, self._dg_as_convT(conv), kind=L.CONVT_4X4_S2, kh=k, kw=k, cin=cout, cout=cin)

    def _conv1d_bwd(self, conv, x_b, g_b, g_for_bias, cin, cout):
        """nn.Conv1d(cin, cout, 1) used as a per-position linear layer (unet_openai.py:322-324)."""
        w = conv.weight
        dwpk = ops.conv_wgrad(x_b, g_b, kind=L.CONV_S1, kh=1, kw=1, cin=cin, cout=cout)
        self._grad(w, ops.unpack_linear_wgrad(dwpk, w.detach().reshape(cout, cin)).view_as(w))
        self._grad(conv.bias, ops.colsum(g_for_bias, cout))
        wt = self.m._cached((id(conv), "dg"), (w,), lambda: ops.pack_weight(w.detach().reshape(cout, cin).contiguous(),
                                                                            1, cin, cout, 0, 1, cin))
        return ops.conv_igemm(g_b, wt, kind=L.CONV_S1, kh=1, kw=1, cin=cout, cout=cin)

    def _linear_bwd(self, lin, x_b, g_b, g_f32, cin, cout, want_dx=True):
        dwpk = ops.conv_wgrad(x_b, g_b, kind=L.CONV_S1, kh=1, kw=1, cin=cin, cout=cout)
        self._grad(lin.weight, ops.unpack_linear_wgrad(dwpk, lin.weight))
        self._grad(lin.bias, ops.colsum(g_f32, cout))
        if not want_dx:
            return None
        return ops.conv_igemm(g_b, self._dg_linear(lin), kind=L.CONV_S1, kh=1, kw=1, cin=cout, cout=cin)

    # ------------------------------------------------------------------ ResBlock (unet_openai.py:291-305)
    def res_block(self, blk, x: _Node, cond, dcond, off, dst) -> _Node:
        m = self.m
        c_in, c_out = blk.channels, blk.out_channels
        gn1, conv1 = blk.in_layers[0], blk.in_layers[2]
        gn2, conv2 = blk.out_layers[0], blk.out_layers[3]
        a, st1 = self._gn32(x.f32, c_in, gn1, L.ACT_SILU)
        ssn = blk.use_scale_shift_norm
        if ssn:   # out_norm(h) * (1 + scale) + shift -> SiLU (unet_openai.py:296-300)
            sc, sh = cond[:, 0, 0, off:off + c_out], cond[:, 0, 0, off + c_out:off + 2 * c_out]
            h = ops.conv_igemm(a, m._w_conv(conv1), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_out, bias=conv1.bias)
            st2 = torch.zeros((h.shape[0], gn2.num_groups, 2), dtype=torch.float64, device=h.device)
            ops.group_stats(h, c_out, gn2.num_groups, st2)
            a2 = torch.empty((*h.shape[:3], pad8(c_out)), dtype=torch.bfloat16, device=h.device)
            ops.groupnorm_apply(h, c_out, st2, gn2.weight, gn2.bias, groups=gn2.num_groups, act=L.ACT_SILU, out=a2,
                                eps=gn2.eps, mod_scale=sc, mod_shift=sh)
        else:
            h = ops.conv_igemm(a, m._w_conv(conv1), kind=L.CONV_S1, kh=3, kw=3, cin=c_in, cout=c_out, bias=conv1.bias,
                               rowbias=cond[:, :, :, off:off + c_out])
            a2, st2 = self._gn32(h, c_out, gn2, L.ACT_SILU)
        drop = self.drop
        p_drop = float(blk.dropout) if drop is not None else 0.0
        layer_id = m._res_index[id(blk)]
        if p_drop > 0:
            ops.dropout_(a2, c_out, p_drop, drop[0], layer_id, drop[1])
        sk = blk.skip_connection
        has_skip = isinstance(sk, nn.Conv2d)
        if has_skip:
            ksk = sk.kernel_size[0]
            res = ops.conv_igemm(x.bf16, m._w_conv(sk), kind=L.CONV_S1, kh=ksk, kw=ksk, cin=c_in, cout=c_out,
                                 bias=sk.bias)
        else:
            res = x.f32
        of, ob = dst(c_out, a2.shape)
        ops.conv_igemm(a2, m._w_conv(conv2), kind=L.CONV_S1, kh=3, kw=3, cin=c_out, cout=c_out, bias=conv2.bias,
                       residual=res, out=of, out2=ob)
        out = _Node(c_out, f32=of, bf16=ob)
        x_b, x_f = x.bf16, x.f32

        def bwd():
            g = out.g
            _, g_b = ops.add(g, None, c_out, want_bf16=True)
            da2 = self._conv2d_bwd(conv2, a2, g_b, g, c_out, c_out, 3)
            if p_drop > 0:
                ops.dropout_(da2, c_out, p_drop, drop[0], layer_id, drop[1])
            if ssn:
                # the normalised tensor n = out_norm(h) is recomputed (fp32), the modulation + SiLU back-propagate in
                # one kernel (dn, per-sample dscale / dshift), then the plain GroupNorm32 backward
                nrm = torch.empty((*h.shape[:3], pad8(c_out)), dtype=torch.float32, device=h.device)
                ops.groupnorm_apply(h, c_out, st2, gn2.weight, gn2.bias, groups=gn2.num_groups, act=L.ACT_NONE,
                                    out_f32=nrm, eps=gn2.eps)
                dn = ops.scale_shift_bwd(nrm, da2, c_out, sc, sh, L.ACT_SILU, dcond[:, 0, 0, off:off + c_out],
                                         dcond[:, 0, 0, off + c_out:off + 2 * c_out])
                dh, dh_b = self._gn32_bwd(gn2, h, dn, c_out, st2, L.ACT_NONE, want_bf16=True)
            else:
                dh, dh_b = self._gn32_bwd(gn2, h, da2, c_out, st2, L.ACT_SILU, want_bf16=True)
                ops.colsum_per_sample(dh, c_out, dcond[:, 0, 0, off:off + c_out])   # d emb_out = sum over the pixels
            da = self._conv2d_bwd(conv1, a, dh_b, dh, c_in, c_out, 3)
            dskip = self._conv2d_bwd(sk, x_b, g_b, g, c_in, c_out, ksk) if has_skip else g
            dx, _ = self._gn32_bwd(gn1, x_f, da, c_in, st1, L.ACT_SILU, addend=dskip)
            _acc(x, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ AttentionBlock (unet_openai.py:329-358)
    def attention(self, blk, x: _Node, dst) -> _Node:
        m = self.m
        c = blk.channels
        a, st = self._gn32(x.f32, c, blk.norm, L.ACT_NONE)
        qkv = ops.conv_igemm(a, m._w_conv1d(blk.qkv), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=3 * c, bias=blk.qkv.bias)
        dh = c // blk.num_heads
        scale = 1.0 / math.sqrt(dh)
        o = ops.softmax_attn(qkv, blk.num_heads, dh, 0, dh, 2 * dh, 3 * dh, scale)
        of, ob = dst(c, a.shape)
        ops.conv_igemm(o, m._w_conv1d(blk.proj_out), kind=L.CONV_S1, kh=1, kw=1, cin=c, cout=c, bias=blk.proj_out.bias,
                       residual=x.f32, out=of, out2=ob)
        out = _Node(c, f32=of, bf16=ob)
        x_f = x.f32

        def bwd():
            g = out.g
            _, g_b = ops.add(g, None, c, want_bf16=True)
            do = self._conv1d_bwd(blk.proj_out, o, g_b, g, c, c)
            dqkv_b = ops.softmax_attn_bwd(qkv, do, blk.num_heads, dh, 0, dh, 2 * dh, 3 * dh, scale, 3 * c)
            da = self._conv1d_bwd(blk.qkv, a, dqkv_b, dqkv_b, c, 3 * c)
            dx, _ = self._gn32_bwd(blk.norm, x_f, da, c, st, L.ACT_NONE, addend=g)
            _acc(x, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ whole network
    def forward(self, x, timesteps, z, y=None):
        from .unet_openai import AttentionBlock, ResBlock
        m = self.m
        b, mch, hh, ww = x.shape
        dev = x.device
        ted, mc = m.time_embed_dim, m.model_channels
        with_z = z is not None
        labels = y   # class labels (UNetModel(num_classes=K)); `y` is rebound to the network output further down
        self.drop = m._dropout_draw() if (m.training and m.dropout > 0) else None

        # ---- embedding path (pre-activations kept for the SiLU backward)
        te = ops.time_embed(timesteps, mc, 1)
        width = (2 if with_z else 1) * ted
        hcat = torch.empty((b, 1, 1, width), dtype=torch.bfloat16, device=dev)
        hpre = torch.empty((b, 1, 1, width), dtype=torch.bfloat16, device=dev)
        l0 = m.time_embed[0]
        ops.conv_igemm(te, m._w_linear(l0), kind=L.CONV_S1, kh=1, kw=1, cin=mc, cout=ted, bias=l0.bias, act=L.ACT_SILU,
                       out=hcat[..., :ted], out2=hpre[..., :ted], out2_preact=True)
        zb = None
        if with_z:
            zb, _ = ops.nchw_to_nhwc(z.view(b, -1, 1, 1))
            p0 = m.proj[0]
            ops.conv_igemm(zb, m._w_linear(p0), kind=L.CONV_S1, kh=1, kw=1, cin=z.shape[1], cout=ted, bias=p0.bias,
                           act=L.ACT_SILU, out=hcat[..., ted:], out2=hpre[..., ted:], out2_preact=True)
        w2, b2 = m._w_emb2(with_z)
        epre = torch.empty((b, 1, 1, pad8(ted)), dtype=torch.bfloat16, device=dev)
        emb_act = ops.conv_igemm(hcat, w2, kind=L.CONV_S1, kh=1, kw=1, cin=width, cout=ted, bias=b2, act=L.ACT_SILU,
                                 out_dtype=torch.bfloat16, out2=epre, out2_preact=True,
                                 rowbias=m._label_rows(labels, b) if labels is not None else None)
        wc, bc, offs, total = m._w_cond()
        cond = ops.conv_igemm(emb_act, wc, kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=total, bias=bc)
        dcond = torch.zeros_like(cond)

        def bwd_emb():
            _, dc_b = ops.add(dcond, None, total, want_bf16=True)
            dbc = ops.colsum(dcond, total)
            dwc = ops.conv_wgrad(emb_act, dc_b, kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=total)
            for blk in m._res_blocks:
                o, lin = offs[id(blk)], blk.emb_layers[1]
                self._grad(lin.weight, ops.unpack_linear_wgrad(dwc[:, o:o + blk.cond_channels], lin.weight))
                self._grad(lin.bias, dbc[o:o + blk.cond_channels].clone())
            params = tuple(blk.emb_layers[1].weight for blk in m._res_blocks)

            def build_t():  # W_cat^T: [1][ted][sum(C_out)]
                wcat = torch.cat([blk.emb_layers[1].weight.detach() for blk in m._res_blocks], dim=0).contiguous()
                return ops.pack_weight(wcat, 1, ted, total, 0, 1, ted)

            de_act = ops.conv_igemm(dc_b, m._cached("cond_dg", params, build_t), kind=L.CONV_S1, kh=1, kw=1, cin=total,
                                    cout=ted)
            de_f, de_b = ops.act_bwd(de_act, epre, ted, L.ACT_SILU, want_f32=True, want_bf16=True)
            lt2 = m.time_embed[2]
            db2 = ops.colsum(de_f, ted)
            if labels is not None:   # d label_emb.weight[k] = sum of d emb over the samples labelled k (row scatter: torch)
                dlab = torch.zeros_like(m.label_emb.weight, dtype=torch.float32)
                dlab.index_add_(0, labels.long(), de_f.view(b, -1)[:, :ted])
                self._grad(m.label_emb.weight, dlab)
            dwt = ops.conv_wgrad(hcat[..., :ted], de_b, kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=ted)
            self._grad(lt2.weight, ops.unpack_linear_wgrad(dwt, lt2.weight))
            self._grad(lt2.bias, db2)
            dt1 = ops.conv_igemm(de_b, self._dg_linear(lt2), kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=ted)
            d1f, d1b = ops.act_bwd(dt1, hpre[..., :ted], ted, L.ACT_SILU, want_f32=True, want_bf16=True)
            self._linear_bwd(l0, te, d1b, d1f, mc, ted, want_dx=False)
            if with_z:
                lz2, lz0 = m.proj[2], m.proj[0]
                dwz = ops.conv_wgrad(hcat[..., ted:], de_b, kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=ted)
                self._grad(lz2.weight, ops.unpack_linear_wgrad(dwz, lz2.weight))
                self._grad(lz2.bias, db2.clone())
                dz1 = ops.conv_igemm(de_b, self._dg_linear(lz2), kind=L.CONV_S1, kh=1, kw=1, cin=ted, cout=ted)
                dzf, dzb = ops.act_bwd(dz1, hpre[..., ted:], ted, L.ACT_SILU, want_f32=True, want_bf16=True)
                self._linear_bwd(lz0, zb, dzb, dzf, z.shape[1], ted, want_dx=False)

        self.tape.append(bwd_emb)  # runs LAST in the backward (the tape is replayed in reverse)

        # ---- concat buffers (same plan as the inference path) + their gradient split
        n_out = len(m.output_blocks)
        cats = [None] * n_out
        skip_nodes = [None] * n_out   # producer node of the second channel range of cat j
        h_nodes = [None] * n_out      # producer node of the first channel range of cat j

        def cat_for(j, shape_bhw):
            if cats[j] is None:
                ch_h, ch_s = m._cat_plan[j]
                bb, h_, w_ = shape_bhw
                cats[j] = (torch.empty((bb, h_, w_, ch_h + ch_s), dtype=torch.float32, device=dev),
                           torch.empty((bb, h_, w_, ch_h + ch_s), dtype=torch.bfloat16, device=dev))
            return cats[j]

        def skip_dst(i):
            j = n_out - 1 - i

            def dst(c, shape):
                cf, cb = cat_for(j, shape[:3])
                ch_h = m._cat_plan[j][0]
                return cf[..., ch_h:ch_h + c], cb[..., ch_h:ch_h + c]
            return dst

        def h_dst(j):
            def dst(c, shape):
                cf, cb = cat_for(j, shape[:3])
                return cf[..., :c], cb[..., :c]
            return dst

        def plain_dst(c, shape):
            return (torch.empty((*shape[:3], pad8(c)), dtype=torch.float32, device=dev),
                    torch.empty((*shape[:3], pad8(c)), dtype=torch.bfloat16, device=dev))

        def downsample(layer, cur, dst):
            h_, w_ = cur.bf16.shape[1:3]
            c = cur.c
            of, ob = dst(c, (b, h_ // 2, w_ // 2))
            xin = cur
            ops.conv_igemm(cur.bf16, m._w_conv(layer.op), kind=L.CONV_S2, kh=3, kw=3, cin=c, cout=c, bias=layer.op.bias,
                           out=of, out2=ob)
            out = _Node(c, f32=of, bf16=ob)

            def bwd():
                g = out.g
                _, g_b = ops.add(g, None, c, want_bf16=True)
                _acc(xin, self._conv2d_bwd(layer.op, xin.bf16, g_b, g, c, c, 3, kind=L.CONV_S2))

            self.tape.append(bwd)
            return out

        def upsample(layer, cur, dst):
            c = cur.c
            xin = cur
            up = ops.upsample_nearest2x(cur.bf16, c)
            of, ob = dst(c, up.shape)
            ops.conv_igemm(up, m._w_conv(layer.conv), kind=L.CONV_S1, kh=3, kw=3, cin=c, cout=c, bias=layer.conv.bias,
                           out=of, out2=ob)
            out = _Node(c, f32=of, bf16=ob)

            def bwd():
                g = out.g
                _, g_b = ops.add(g, None, c, want_bf16=True)
                dup = self._conv2d_bwd(layer.conv, up, g_b, g, c, c, 3)
                _acc(xin, ops.upsample_nearest2x_bwd(dup, c))

            self.tape.append(bwd)
            return out

        def run_layers(layers, cur, last_dst):
            for li, layer in enumerate(layers):
                dst = last_dst if li == len(layers) - 1 else plain_dst
                if isinstance(layer, ResBlock):
                    cur = self.res_block(layer, cur, cond, dcond, offs[id(layer)], dst)
                elif isinstance(layer, AttentionBlock):
                    cur = self.attention(layer, cur, dst)
                elif hasattr(layer, "op"):
                    cur = downsample(layer, cur, dst)
                else:
                    cur = upsample(layer, cur, dst)
            return cur

        # ---- input blocks
        xb, _ = ops.nchw_to_nhwc(x)
        stem = m.input_blocks[0][0]
        of, ob = skip_dst(0)(mc, (b, hh, ww))
        ops.conv_igemm(xb, m._w_conv(stem), kind=L.CONV_S1, kh=3, kw=3, cin=mch, cout=mc, bias=stem.bias, out=of, out2=ob)
        cur = _Node(mc, f32=of, bf16=ob)
        stem_node = cur

        def bwd_stem():
            g = stem_node.g
            _, g_b = ops.add(g, None, mc, want_bf16=True)
            self._conv2d_bwd(stem, xb, g_b, g, mch, mc, 3, want_dx=False)

        self.tape.append(bwd_stem)
        skip_nodes[n_out - 1] = cur
        for i, block in enumerate(m.input_blocks):
            if i == 0:
                continue
            cur = run_layers(list(block), cur, skip_dst(i))
            skip_nodes[n_out - 1 - i] = cur

        # ---- middle block
        mb = list(m.middle_block)
        cur = self.res_block(mb[0], cur, cond, dcond, offs[id(mb[0])], plain_dst)
        cur = self.attention(mb[1], cur, plain_dst)
        cur = self.res_block(mb[2], cur, cond, dcond, offs[id(mb[2])], h_dst(0))
        h_nodes[0] = cur

        # ---- output blocks
        for j, block in enumerate(m.output_blocks):
            cf, cb = cats[j]
            ch_h = m._cat_plan[j][0]
            cat_node = _Node(cf.shape[-1], f32=cf, bf16=cb)

            def bwd_cat(cat_node=cat_node, first=h_nodes[j], skip=skip_nodes[j], ch_h=ch_h):
                g = cat_node.g
                _acc(first, g[..., :ch_h])
                _acc(skip, g[..., ch_h:])

            self.tape.append(bwd_cat)
            cur = run_layers(list(block), cat_node, h_dst(j + 1) if j + 1 < n_out else plain_dst)
            if j + 1 < n_out:
                h_nodes[j + 1] = cur

        # ---- out: GroupNorm32 -> SiLU -> conv3x3 -> NCHW fp32
        gno, oc = m.out[0], m.out[2]
        a, st = self._gn32(cur.f32, cur.c, gno, L.ACT_SILU)
        y = ops.conv_igemm(a, m._w_conv(oc), kind=L.CONV_S1, kh=3, kw=3, cin=cur.c, cout=m.out_channels, bias=oc.bias,
                           nchw=True)
        last_node, c_last = cur, cur.c

        def bwd_last(dout):
            d_b, d_f = ops.nchw_to_nhwc(dout, want_f32=True)
            da = self._conv2d_bwd(oc, a, d_b, d_f, c_last, m.out_channels, 3)
            dx, _ = self._gn32_bwd(gno, last_node.f32, da, c_last, st, L.ACT_SILU)
            last_node.g = dx

        self.bwd_last = bwd_last
        return y


class UNetModelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, timesteps, z, y, *params):
        if x.requires_grad or timesteps.requires_grad or (z is not None and z.requires_grad):
            raise L.SbmError("UNetModel: gradients w.r.t. the input latent / timesteps / conditioning code z are not "
                             "implemented on the B200 path (detach them, or differentiate the parameters only)")
        plan = _OPlan(model)
        with torch.no_grad():
            out = plan.forward(x.contiguous().float(), timesteps.contiguous().float(),
                               None if z is None else z.contiguous().float(), y)
        ctx.plan = plan
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        if plan is None:
            raise L.SbmError("UNetModel: the backward tape was already consumed (a second backward through the same "
                             "forward is not supported: run the forward again)")
        with torch.no_grad():
            plan.backward(dout)
        grads = tuple(plan.pg.get(p) for p in ctx.params)
        ctx.plan = None
        return (None, None, None, None, None) + grads


def unet_openai_forward_train(model, x, timesteps, z=None, y=None):
    params = tuple(p for p in model.parameters() if p.requires_grad)
    return UNetModelFn.apply(model, x, timesteps, z, y, *params)
