"""Operator-level host wrappers: torch tensors in, C-ABI calls out.

Each function is a thin marshalling layer over one `sbm_*` entry point of
libsbmae_b200 (include/sbmae_b200.h); no arithmetic happens in Python.
Activations are channels-last `[B, H, W, ld]` tensors (ld = channels rounded up to 8).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


_ITEMSIZE = {torch.float32: 4, torch.float64: 8, torch.bfloat16: 2, torch.int32: 4, torch.int64: 8}


class _ZeroArena:
    """The backward pass needs ~400 small zero-initialised accumulators (split-K weight gradients, bias / GroupNorm
    parameter gradients, reduction scratch).  Instead of one fill kernel each, a backward pass takes them from ONE
    freshly allocated zeroed buffer (sized from the previous pass), i.e. one fill launch.  The buffer is new every
    pass, so gradients that alias it stay valid for as long as they are referenced."""

    def __init__(self):
        self.buf, self.off, self.used = None, 0, 0

    def begin(self, size_hint: int, device) -> None:
        self.buf = torch.zeros(size_hint, dtype=torch.float32, device=device) if size_hint > 0 else None
        self.off, self.used = 0, 0

    def end(self) -> int:
        used, self.buf = self.used, None
        return used

    def take(self, shape, dtype, device):
        n = 1
        for d in shape:
            n *= d
        words = (n * _ITEMSIZE[dtype] + 3) // 4
        words = (words + 3) // 4 * 4  # 16-byte aligned slices
        self.used += words
        if self.buf is None or self.buf.device != device or self.off + words > self.buf.numel():
            return torch.zeros(shape, dtype=dtype, device=device)
        raw = self.buf[self.off:self.off + words]
        self.off += words
        return raw.view(dtype)[:n].view(shape)


_arena: _ZeroArena | None = None


def arena_begin(size_hint: int, device) -> None:
    global _arena
    _arena = _ZeroArena()
    _arena.begin(size_hint, device)


def arena_end() -> int:
    global _arena
    used = _arena.end() if _arena is not None else 0
    _arena = None
    return used


def _zeros(shape, dtype, device):
    shape = tuple(shape) if isinstance(shape, (tuple, list)) else (shape,)
    if _arena is not None:
        return _arena.take(shape, dtype, torch.device(device))
    return torch.zeros(shape, dtype=dtype, device=device)


def pack_weight(w: torch.Tensor, taps: int, rows: int, cols: int, s_tap: int, s_row: int, s_col: int,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 weight (any strided view described by element strides) -> bf16 [taps, rows, pad8(cols)]."""
    assert w.dtype == torch.float32 and w.is_cuda
    cols_pad = pad8(cols)
    if out is None:
        out = torch.empty((taps, rows, cols_pad), dtype=torch.bfloat16, device=w.device)
    L.check(L.lib().sbm_pack_weight_bf16(L.ptr(w), L.ptr(out), C.c_int32(taps), C.c_int32(rows), C.c_int32(cols),
                                         C.c_int32(cols_pad), C.c_int64(s_tap), C.c_int64(s_row), C.c_int64(s_col),
                                         L.stream_ptr()), "sbm_pack_weight_bf16")
    # how to redo this pack in place (pack_weights_multi): `w` keeps the source storage alive / addressable
    out._pack_spec = (w, taps, rows, cols, cols_pad, s_tap, s_row, s_col)
    return out


_multi_tables: dict = {}


def pack_weights_multi(packs) -> None:
    """Re-run the packs of `packs` (tensors returned by pack_weight, refreshed IN PLACE from their current source
    values) as ONE kernel launch.  The device descriptor table is cached per set of (source, destination) pointers."""
    if not packs:
        return
    # full pack geometry in the key: after a model is freed the allocator may hand the same addresses to tensors of
    # another shape; bounded (a training run keeps re-using ONE table per model)
    key = tuple((t._pack_spec[0].data_ptr(), t.data_ptr()) + tuple(t._pack_spec[1:]) for t in packs)
    hit = _multi_tables.get(key)
    if hit is None and len(_multi_tables) >= 16:
        _multi_tables.pop(next(iter(_multi_tables)))
    if hit is None:
        arr = (L.PackDesc * len(packs))()
        blocks, max_taps = 0, 1
        for i, t in enumerate(packs):
            w, taps, rows, cols, cols_pad, s_tap, s_row, s_col = t._pack_spec
            tiles_c = (cols_pad + 31) // 32
            arr[i] = L.PackDesc(w.data_ptr(), t.data_ptr(), taps, rows, cols, cols_pad, s_tap, s_row, s_col, tiles_c,
                                blocks, ((1 << 32) + taps - 1) // taps if taps > 1 else 0, 0)
            blocks += tiles_c * ((rows + 31) // 32)
            max_taps = max(max_taps, taps)
        dev = packs[0].device
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        hit = (table, len(packs), blocks, max_taps)
        _multi_tables[key] = hit
    table, n, blocks, max_taps = hit
    L.check(L.lib().sbm_pack_weights_multi(L.ptr(table), C.c_int32(n), C.c_int32(blocks), C.c_int32(max_taps),
                                           L.stream_ptr()), "sbm_pack_weights_multi")


def pack_conv2d_weight(w: torch.Tensor, out=None) -> torch.Tensor:
    """nn.Conv2d weight [O, I, KH, KW] -> [KH*KW, O, pad8(I)] bf16."""
    w = w.detach().contiguous()
    o, i, kh, kw = w.shape
    return pack_weight(w, kh * kw, o, i, 1, i * kh * kw, kh * kw, out)


def pack_convT2d_weight(w: torch.Tensor, out=None) -> torch.Tensor:
    """nn.ConvTranspose2d weight [I, O, KH, KW] -> [KH*KW, O, pad8(I)] bf16."""
    w = w.detach().contiguous()
    i, o, kh, kw = w.shape
    return pack_weight(w, kh * kw, o, i, 1, kh * kw, o * kh * kw, out)


def fold_groupnorm_conv(w: torch.Tensor, bias, gamma: torch.Tensor, beta: torch.Tensor):
    """nn.GroupNorm(1, I) followed by nn.Conv2d(I, O, k) -> (bf16 [k*k, O, pad8(I)] weights carrying gamma,
    fp32 [2, 16, O] border-class tables carrying mean-correction sums, beta and the conv bias)."""
    w = w.detach().contiguous()
    o, i, kh, kw = w.shape
    cols_pad = pad8(i)
    wpk = torch.empty((kh * kw, o, cols_pad), dtype=torch.bfloat16, device=w.device)
    tab = torch.empty((2, 16, o), dtype=torch.float32, device=w.device)
    L.check(L.lib().sbm_conv_fold_groupnorm(L.ptr(w), L.ptr(wpk), L.ptr(tab), C.c_int32(kh), C.c_int32(kw), C.c_int32(o),
                                            C.c_int32(i), C.c_int32(cols_pad), C.c_int64(1), C.c_int64(i * kh * kw),
                                            C.c_int64(kh * kw), L.ptr(gamma.detach()), L.ptr(beta.detach()),
                                            L.ptr(bias.detach() if bias is not None else None), L.stream_ptr()),
            "sbm_conv_fold_groupnorm")
    return wpk, tab


def pack_linear_weight(w: torch.Tensor, out=None) -> torch.Tensor:
    """nn.Linear weight [O, I] -> [1, O, pad8(I)] bf16."""
    w = w.detach().contiguous()
    o, i = w.shape
    return pack_weight(w, 1, o, i, 0, i, 1, out)


_splitk_ws: dict = {}


def _splitk_workspace(dev: torch.device, elems: int) -> torch.Tensor:
    """fp32 workspace of a split-K convolution (per-slice slabs, summed by the library's second kernel; no zeroing
    needed).  ONE buffer per device serves every call in stream order; the address is stable, so the calls can be
    replayed from a CUDA graph.  Like the packed-weight caches it assumes that one device's score-net calls are issued
    in order (one stream at a time, or streams that are joined between calls)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _splitk_ws.get(key)
    if ws is None or ws.numel() < elems:
        if torch.cuda.is_current_stream_capturing():
            # growing inside a capture would put the buffer into the graph's private pool: size it beforehand
            raise L.SbmError("split-K workspace too small inside CUDA-graph capture (run the call once eagerly first)")
        ws = _splitk_ws[key] = torch.empty(max(elems, 1 << 24), dtype=torch.float32, device=dev)
    return ws


def conv_igemm(x: torch.Tensor, wpk: torch.Tensor, *, kind: int, kh: int, kw: int, cin: int, cout: int,
               bias: torch.Tensor | None = None, act: int = L.ACT_NONE, residual: torch.Tensor | None = None,
               out: torch.Tensor | None = None, out_dtype: torch.dtype = torch.float32, nchw: bool = False,
               stats: torch.Tensor | None = None, out2: torch.Tensor | None = None,
               out2_preact: bool = False, rowbias: torch.Tensor | None = None, gn_stats: torch.Tensor | None = None,
               gn_tab: torch.Tensor | None = None, gn_eps: float = 1e-5) -> torch.Tensor:
    """x: bf16 [B,H,W,ldx]; returns [B,OH,OW,pad8(cout)] (or fp32 NCHW [B,cout,OH,OW] when nchw)."""
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.stride(3) == 1
    b, h, w, _ = x.shape
    ldx = x.stride(2)
    assert x.stride(1) == w * ldx and x.stride(0) == h * w * ldx, "activation must be dense pixel-major"
    if kind == L.CONV_S1:
        oh, ow = h, w
    elif kind == L.CONV_S2:
        oh, ow = h // 2, w // 2
    else:
        oh, ow = 2 * h, 2 * w
    if out is None:
        if nchw:
            out = torch.empty((b, cout, oh, ow), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((b, oh, ow, pad8(cout)), dtype=out_dtype, device=x.device)
    a = L.ConvArgs()
    a.kind, a.kh, a.kw = kind, kh, kw
    a.batch, a.h, a.w = b, h, w
    a.cin, a.cout = cin, cout
    a.x, a.ldx = x.data_ptr(), ldx
    a.wpk, a.cin_pad = wpk.data_ptr(), wpk.shape[2]
    a.act = act
    a.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        a.residual, a.ldr = residual.data_ptr(), residual.stride(2)
        a.res_dtype = L.BF16 if residual.dtype == torch.bfloat16 else L.F32
    a.out = out.data_ptr()
    a.ldo = 0 if nchw else out.stride(2)
    a.out_dtype = L.BF16 if out.dtype == torch.bfloat16 else L.F32
    a.out_nchw = 1 if nchw else 0
    a.stats = stats.data_ptr() if stats is not None else None
    if out2 is not None:
        a.out2, a.ldo2 = out2.data_ptr(), out2.stride(2)
        a.out2_preact = 1 if out2_preact else 0
    if rowbias is not None:
        a.rowbias, a.ld_rowbias = rowbias.data_ptr(), rowbias.stride(-2)
    if gn_tab is not None:  # GroupNorm(1, cin) of x folded into the weights (fold_groupnorm_conv)
        a.gn_stats, a.gn_tab = gn_stats.data_ptr(), gn_tab.data_ptr()
        a.gn_count, a.gn_eps = float(h * w * cin), gn_eps
    # sub-wave K-long layers (low-resolution levels at small batch): split-K needs an fp32 workspace; the plan lives in
    # the library (one source of truth), the allocation here
    # (cheap host-side screen first: the library only splits layers of a few 256 x 256 tiles -- at most half the SM pairs
    # of the largest part, 128 -- with more than 128 output channels; every other call skips the extra ABI round trip)
    if kind in (L.CONV_S1, L.CONV_S2) and not nchw and cout > 128 and 128 <= b * oh * ow <= 64 * 256:
        a.ld_ws = pad8(cout)
        need = L.lib().sbm_conv_splitk_ws_elems(C.byref(a))
        if need > 0:
            ws = _splitk_workspace(x.device, need)
            a.splitk_ws, a.ws_elems = ws.data_ptr(), ws.numel()
    L.check(L.lib().sbm_conv_igemm(C.byref(a), L.stream_ptr()), "sbm_conv_igemm")
    return out


# ---------------------------------------------------------------------------- score-net operators
def _dt(t: torch.Tensor) -> int:
    return L.BF16 if t.dtype == torch.bfloat16 else L.F32


def stem_im2col(x: torch.Tensor, kh: int, kw: int) -> torch.Tensor:
    """fp32 NCHW -> bf16 [B,H,W,pad8(C*kh*kw)] im2col rows."""
    b, c, h, w = x.shape
    ldk = pad8(c * kh * kw)
    a = torch.empty((b, h, w, ldk), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().sbm_stem_im2col(L.ptr(x), L.ptr(a), C.c_int32(b), C.c_int32(c), C.c_int32(h), C.c_int32(w),
                                    C.c_int32(kh), C.c_int32(kw), C.c_int32(ldk), L.stream_ptr()), "sbm_stem_im2col")
    return a


def dwconv7(x: torch.Tensor, c: int, w: torch.Tensor, bias, cond, ldc: int, stats,
            out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    b, h, wd, _ = x.shape
    out = torch.empty((b, h, wd, pad8(c)), dtype=out_dtype, device=x.device)
    fn = L.lib().sbm_dwconv7_fwd_bf16 if out_dtype == torch.bfloat16 else L.lib().sbm_dwconv7_fwd
    L.check(fn(L.ptr(x), C.c_int64(x.stride(2)), L.ptr(w), L.ptr(bias), L.ptr(cond), C.c_int64(ldc), L.ptr(out),
               C.c_int64(out.stride(2)), L.ptr(stats), C.c_int32(b), C.c_int32(h), C.c_int32(wd), C.c_int32(c),
               L.stream_ptr()), "sbm_dwconv7_fwd")
    return out


def group_stats(x: torch.Tensor, c: int, groups: int, stats: torch.Tensor) -> None:
    b, h, w, _ = x.shape
    L.check(L.lib().sbm_group_stats(L.ptr(x), C.c_int32(_dt(x)), C.c_int64(x.stride(2)), C.c_int32(b),
                                    C.c_int32(h * w), C.c_int32(c), C.c_int32(groups), L.ptr(stats), L.stream_ptr()),
            "sbm_group_stats")


def groupnorm_apply(x: torch.Tensor, c: int, stats: torch.Tensor, gamma, beta, *, groups: int = 1, act: int = 0,
                    residual=None, out=None, out_f32=None, eps: float = 1e-5, mod_scale=None, mod_shift=None,
                    post_add=None) -> None:
    """y = act(GroupNorm(x) [* (1 + mod_scale[b]) + mod_shift[b]]) [+ post_add[b]] [+ residual]; the per-sample
    operands are fp32 [B, >=c] views (row stride = stride(0))."""
    b, h, w, _ = x.shape
    if mod_scale is not None:
        assert mod_shift is not None and mod_scale.stride(0) == mod_shift.stride(0)
    L.check(L.lib().sbm_groupnorm_apply_mod(
        L.ptr(x), C.c_int32(_dt(x)), C.c_int64(x.stride(2)), L.ptr(stats), L.ptr(gamma), L.ptr(beta),
        L.ptr(residual), C.c_int64(residual.stride(2) if residual is not None else 0),
        L.ptr(out), C.c_int32(_dt(out) if out is not None else L.BF16),
        C.c_int64(out.stride(2) if out is not None else 4),
        L.ptr(out_f32), C.c_int64(out_f32.stride(2) if out_f32 is not None else 0),
        C.c_int32(b), C.c_int32(h * w), C.c_int32(c), C.c_int32(groups), C.c_float(eps), C.c_int32(act),
        L.ptr(mod_scale), L.ptr(mod_shift), C.c_int64(mod_scale.stride(0) if mod_scale is not None else 0),
        L.ptr(post_add), C.c_int64(post_add.stride(0) if post_add is not None else 0),
        L.stream_ptr()), "sbm_groupnorm_apply_mod")


def dropout_(x: torch.Tensor, c: int, p: float, seed: int, draw: int, draw_dev=None) -> torch.Tensor:
    """nn.Dropout(p) in training mode, IN PLACE on a channels-last [B, H, W, ld] tensor (fp32 / bf16); the mask is a
    function of (seed, draw + *draw_dev, element), so the same call on the gradient of the output is the backward."""
    b, h, w, _ = x.shape
    r = L.Rng(seed & 0xFFFFFFFFFFFFFFFF, draw, 0, draw_dev.data_ptr() if draw_dev is not None else None)
    L.check(L.lib().sbm_dropout(L.ptr(x), C.c_int32(_dt(x)), C.c_int64(x.stride(2)), C.c_int64(b * h * w),
                                C.c_int32(c), C.c_float(p), C.byref(r), L.stream_ptr()), "sbm_dropout")
    return x


def time_embed(t: torch.Tensor, dim: int, mode: int) -> torch.Tensor:
    b = t.shape[0]
    ld = pad8(dim)
    out = torch.empty((b, 1, 1, ld), dtype=torch.bfloat16, device=t.device)
    L.check(L.lib().sbm_time_embed(L.ptr(t), L.ptr(out), None, C.c_int32(b), C.c_int32(dim), C.c_int32(ld),
                                   C.c_int32(mode), L.stream_ptr()), "sbm_time_embed")
    return out


def linear_attn(qkv: torch.Tensor, heads: int, scale: float) -> torch.Tensor:
    b, h, w, _ = qkv.shape
    out = torch.empty((b, h, w, heads * 32), dtype=torch.bfloat16, device=qkv.device)
    if qkv.dtype == torch.bfloat16:
        L.check(L.lib().sbm_linear_attn_fwd_bf16(L.ptr(qkv), C.c_int64(qkv.stride(2)), L.ptr(out),
                                                 C.c_int64(out.stride(2)), C.c_int32(b), C.c_int32(h * w),
                                                 C.c_int32(heads), C.c_float(scale), L.stream_ptr()),
                "sbm_linear_attn_fwd_bf16")
        return out
    L.check(L.lib().sbm_linear_attn_fwd(L.ptr(qkv), C.c_int64(qkv.stride(2)), L.ptr(out), C.c_int64(out.stride(2)),
                                        C.c_int32(b), C.c_int32(h * w), C.c_int32(heads), C.c_float(scale),
                                        L.stream_ptr()), "sbm_linear_attn_fwd")
    return out


def softmax_attn(qkv: torch.Tensor, heads: int, dh: int, q_off: int, k_off: int, v_off: int, head_stride: int,
                 scale: float) -> torch.Tensor:
    b, h, w, _ = qkv.shape
    out = torch.empty((b, h, w, pad8(heads * dh)), dtype=torch.bfloat16, device=qkv.device)
    L.check(L.lib().sbm_softmax_attn_fwd(L.ptr(qkv), C.c_int64(qkv.stride(2)), L.ptr(out), C.c_int64(out.stride(2)),
                                         C.c_int32(b), C.c_int32(h * w), C.c_int32(heads), C.c_int32(dh),
                                         C.c_int32(q_off), C.c_int32(k_off), C.c_int32(v_off), C.c_int32(head_stride),
                                         C.c_float(scale), L.stream_ptr()), "sbm_softmax_attn_fwd")
    return out


# ---------------------------------------------------------------------------- backward operators
def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, *, kind: int, kh: int, kw: int, cin: int, cout: int) -> torch.Tensor:
    """Packed fp32 weight gradient [kh*kw, cout, pad8(cin)] of conv_igemm(x, ...) given dy (both bf16 channels-last)."""
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16
    b, h, w, _ = x.shape
    dwpk = _zeros((kh * kw, cout, pad8(cin)), torch.float32, x.device)
    a = L.WgradArgs()
    a.kind, a.kh, a.kw = kind, kh, kw
    a.batch, a.h, a.w = b, h, w
    a.cin, a.cout = cin, cout
    a.x, a.ldx = x.data_ptr(), x.stride(2)
    a.dy, a.lddy = dy.data_ptr(), dy.stride(2)
    a.dwpk, a.cin_pad = dwpk.data_ptr(), dwpk.shape[2]
    L.check(L.lib().sbm_conv_wgrad(C.byref(a), L.stream_ptr()), "sbm_conv_wgrad")
    return dwpk


class _UnpackBatch:
    """Weight-gradient unpacks of one backward pass, deferred and issued as ONE `sbm_unpack_wgrad_multi` launch per
    flush (end of the pass; data-parallel training: before a gradient bucket goes on the wire).  The descriptor table
    of flush k lives in its own pinned staging buffer + device buffer, allocated outside CUDA-graph capture and refreshed
    in place (a captured copy node re-reads the pinned buffer at every replay, like FusedAdam's tables)."""

    def __init__(self):
        self.items = []          # (dwpk, out, taps, rows, cols, cols_pad, s_tap, s_row, s_col): tensors kept alive
        self.flushes = 0
        self.stages = {}         # flush index -> [pinned bytes, device bytes, last signature]

    def begin(self):
        self.items, self.flushes = [], 0

    def add(self, *item):
        self.items.append(item)

    def flush(self):
        if not self.items:
            return
        items, self.items = self.items, []
        k, self.flushes = self.flushes, self.flushes + 1
        sig = tuple((it[0].data_ptr(), it[1].data_ptr()) + it[2:] for it in items)
        nbytes = C.sizeof(L.PackDesc) * len(items)
        st = self.stages.get(k)
        dev = items[0][0].device
        if st is None or st[0].numel() < nbytes or st[1].device != dev:
            cap = max(nbytes, C.sizeof(L.PackDesc) * 256)
            st = [torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(cap, dtype=torch.uint8, device=dev), None]
            self.stages[k] = st
        blocks, max_taps = 0, 1
        arr = (L.PackDesc * len(items))()
        for i, (dwpk, out, taps, rows, cols, cols_pad, s_tap, s_row, s_col) in enumerate(items):
            tiles_c = (cols + 31) // 32
            arr[i] = L.PackDesc(dwpk.data_ptr(), out.data_ptr(), taps, rows, cols, cols_pad, s_tap, s_row, s_col, tiles_c,
                                blocks, ((1 << 32) + taps - 1) // taps if taps > 1 else 0, 0)
            blocks += tiles_c * ((rows + 31) // 32)
            max_taps = max(max_taps, taps)
        if st[2] != sig:
            st[0][:nbytes].copy_(torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8))
            st[2] = sig
        st[1][:nbytes].copy_(st[0][:nbytes], non_blocking=True)
        L.check(L.lib().sbm_unpack_wgrad_multi(L.ptr(st[1]), C.c_int32(len(items)), C.c_int32(blocks),
                                               C.c_int32(max_taps), L.stream_ptr()), "sbm_unpack_wgrad_multi")


_unpack_batch = _UnpackBatch()
_unpack_active = False


def unpack_begin() -> None:
    """Defer the large weight-gradient unpacks that follow until `unpack_flush()`."""
    global _unpack_active
    _unpack_batch.begin()
    _unpack_active = True


def unpack_flush() -> None:
    _unpack_batch.flush()


def unpack_end() -> None:
    global _unpack_active
    _unpack_batch.flush()
    _unpack_active = False


def unpack_wgrad(dwpk: torch.Tensor, like: torch.Tensor, cols: int, s_tap: int, s_row: int, s_col: int,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """Packed gradient -> a tensor shaped like the parameter `like` (strides as in pack_weight).  `out`: a contiguous
    fp32 destination of that shape (data-parallel training: the parameter's slot in the flat bucket buffer).  Inside
    `unpack_begin()` ... `unpack_end()` the tiled unpacks are batched into one launch per flush: the returned tensor is
    filled when the batch is flushed."""
    taps, rows, cols_pad = dwpk.shape
    if out is None:
        out = torch.empty_like(like, dtype=torch.float32, memory_format=torch.contiguous_format)
    if _unpack_active and rows * cols >= 4096 and taps <= 16 and dwpk.is_contiguous():
        _unpack_batch.add(dwpk, out, taps, rows, cols, cols_pad, s_tap, s_row, s_col)
        return out
    L.check(L.lib().sbm_unpack_wgrad(L.ptr(dwpk), L.ptr(out), C.c_int32(taps), C.c_int32(rows), C.c_int32(cols),
                                     C.c_int32(cols_pad), C.c_int64(s_tap), C.c_int64(s_row), C.c_int64(s_col),
                                     L.stream_ptr()), "sbm_unpack_wgrad")
    return out


def unpack_conv2d_wgrad(dwpk, weight, out=None):
    o, i, kh, kw = weight.shape
    return unpack_wgrad(dwpk, weight, i, 1, i * kh * kw, kh * kw, out)


def unpack_convT2d_wgrad(dwpk, weight, out=None):
    i, o, kh, kw = weight.shape
    return unpack_wgrad(dwpk, weight, i, 1, kh * kw, o * kh * kw, out)


def unpack_linear_wgrad(dwpk, weight, out=None):
    o, i = weight.shape
    return unpack_wgrad(dwpk, weight, i, 0, i, 1, out)


def _rows(t: torch.Tensor) -> int:
    return t.shape[0] * t.shape[1] * t.shape[2]


def colsum(x: torch.Tensor, c: int) -> torch.Tensor:
    out = _zeros(c, torch.float32, x.device)
    L.check(L.lib().sbm_colsum(L.ptr(x), C.c_int32(_dt(x)), C.c_int64(x.stride(2)), C.c_int64(_rows(x)), C.c_int32(c),
                               L.ptr(out), L.stream_ptr()), "sbm_colsum")
    return out


def groupnorm_bwd(x, dy, c, stats, gamma, *, groups=1, in_act=0, addend=None, want_f32=True, want_bf16=False,
                  eps=1e-5, beta=None, out_act=0):
    """-> (dx_f32 | None, dx_bf16 | None, dgamma, dbeta).  out_act: activation applied after the norm (needs beta)."""
    b, h, w, _ = x.shape
    dev = x.device
    bst = _zeros((b, groups, 2), torch.float32, dev)
    dgamma = _zeros(c, torch.float32, dev)
    dbeta = _zeros(c, torch.float32, dev)
    of = torch.empty((b, h, w, pad8(c)), dtype=torch.float32, device=dev) if want_f32 else None
    ob = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    L.check(L.lib().sbm_groupnorm_bwd(
        L.ptr(x), C.c_int32(_dt(x)), C.c_int64(x.stride(2)), L.ptr(dy), C.c_int32(_dt(dy)), C.c_int64(dy.stride(2)),
        L.ptr(stats), L.ptr(gamma), L.ptr(bst), L.ptr(dgamma), L.ptr(dbeta),
        L.ptr(addend), C.c_int64(addend.stride(2) if addend is not None else 0),
        L.ptr(of), C.c_int64(of.stride(2) if of is not None else 0),
        L.ptr(ob), C.c_int64(ob.stride(2) if ob is not None else 0),
        C.c_int32(b), C.c_int32(h * w), C.c_int32(c), C.c_int32(groups), C.c_float(eps), C.c_int32(in_act),
        L.ptr(beta), C.c_int32(out_act), L.stream_ptr()), "sbm_groupnorm_bwd")
    return of, ob, dgamma, dbeta


def dwconv7_bwd_input(dy, c, w, addend=None):
    b, h, wd, _ = dy.shape
    out = torch.empty((b, h, wd, pad8(c)), dtype=torch.float32, device=dy.device)
    L.check(L.lib().sbm_dwconv7_bwd_input(L.ptr(dy), C.c_int64(dy.stride(2)), L.ptr(w), L.ptr(addend),
                                          C.c_int64(addend.stride(2) if addend is not None else 0), L.ptr(out),
                                          C.c_int64(out.stride(2)), C.c_int32(b), C.c_int32(h), C.c_int32(wd),
                                          C.c_int32(c), L.stream_ptr()), "sbm_dwconv7_bwd_input")
    return out


def dwconv7_wgrad(x, dy, c, dcond=None, ldc=0, want_db=True):
    """-> (dw [c,1,7,7], db [c] | None); dcond (a [B, >=c] slice view) is overwritten with sum_p dy."""
    b, h, wd, _ = x.shape
    dw = _zeros((c, 1, 7, 7), torch.float32, x.device)
    db = _zeros(c, torch.float32, x.device) if want_db else None
    L.check(L.lib().sbm_dwconv7_wgrad(L.ptr(x), C.c_int64(x.stride(2)), L.ptr(dy), C.c_int64(dy.stride(2)), L.ptr(dw),
                                      L.ptr(db), L.ptr(dcond), C.c_int64(ldc), C.c_int32(b), C.c_int32(h),
                                      C.c_int32(wd), C.c_int32(c), L.stream_ptr()), "sbm_dwconv7_wgrad")
    return dw, db


def linear_attn_bwd(qkv, dout, heads, scale):
    b, h, w, _ = qkv.shape
    dqkv = torch.empty((b, h, w, 3 * heads * 32), dtype=torch.bfloat16, device=qkv.device)
    L.check(L.lib().sbm_linear_attn_bwd(L.ptr(qkv), C.c_int64(qkv.stride(2)), L.ptr(dout), C.c_int64(dout.stride(2)),
                                        L.ptr(dqkv), C.c_int64(dqkv.stride(2)), C.c_int32(b), C.c_int32(h * w),
                                        C.c_int32(heads), C.c_float(scale), L.stream_ptr()), "sbm_linear_attn_bwd")
    return dqkv


def softmax_attn_bwd(qkv, dout, heads, dh, q_off, k_off, v_off, head_stride, scale, width):
    b, h, w, _ = qkv.shape
    dqkv = torch.zeros((b, h, w, pad8(width)), dtype=torch.bfloat16, device=qkv.device)
    L.check(L.lib().sbm_softmax_attn_bwd(L.ptr(qkv), C.c_int64(qkv.stride(2)), L.ptr(dout), C.c_int64(dout.stride(2)),
                                         L.ptr(dqkv), C.c_int64(dqkv.stride(2)), C.c_int32(b), C.c_int32(h * w),
                                         C.c_int32(heads), C.c_int32(dh), C.c_int32(q_off), C.c_int32(k_off),
                                         C.c_int32(v_off), C.c_int32(head_stride), C.c_float(scale), L.stream_ptr()),
            "sbm_softmax_attn_bwd")
    return dqkv


def act_bwd(dy, pre, c, act, want_f32=False, want_bf16=True):
    b, h, w, _ = dy.shape
    of = torch.empty((b, h, w, pad8(c)), dtype=torch.float32, device=dy.device) if want_f32 else None
    ob = torch.empty((b, h, w, pad8(c)), dtype=torch.bfloat16, device=dy.device) if want_bf16 else None
    L.check(L.lib().sbm_act_bwd(L.ptr(dy), C.c_int64(dy.stride(2)), L.ptr(pre), C.c_int32(_dt(pre)),
                                C.c_int64(pre.stride(2)), L.ptr(of), C.c_int64(of.stride(2) if of is not None else 0),
                                L.ptr(ob), C.c_int64(ob.stride(2) if ob is not None else 0), C.c_int64(_rows(dy)),
                                C.c_int32(c), C.c_int32(act), L.stream_ptr()), "sbm_act_bwd")
    return of, ob


def nchw_to_nhwc(x, want_f32=False):
    b, c, h, w = x.shape
    ob = torch.zeros((b, h, w, pad8(c)), dtype=torch.bfloat16, device=x.device)
    of = torch.zeros((b, h, w, pad8(c)), dtype=torch.float32, device=x.device) if want_f32 else None
    L.check(L.lib().sbm_nchw_to_nhwc(L.ptr(x), L.ptr(ob), C.c_int64(ob.stride(2)), L.ptr(of),
                                     C.c_int64(of.stride(2) if of is not None else 0), C.c_int32(b), C.c_int32(c),
                                     C.c_int32(h * w), L.stream_ptr()), "sbm_nchw_to_nhwc")
    return ob, of


def add(a, b, c, out=None, want_bf16=False):
    """fp32 channels-last a + b (b may be None) -> (out fp32 | None, bf16 copy | None)."""
    bb, h, w, _ = a.shape
    ob = torch.empty((bb, h, w, pad8(c)), dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    L.check(L.lib().sbm_add(L.ptr(a), C.c_int64(a.stride(2)), L.ptr(b), C.c_int64(b.stride(2) if b is not None else 0),
                            L.ptr(out), C.c_int64(out.stride(2) if out is not None else 0), L.ptr(ob),
                            C.c_int64(ob.stride(2) if ob is not None else 0), C.c_int64(_rows(a)), C.c_int32(c),
                            L.stream_ptr()), "sbm_add")
    return out, ob


def upsample_nearest2x(x: torch.Tensor, c: int) -> torch.Tensor:
    b, h, w, ld = x.shape
    out = torch.empty((b, 2 * h, 2 * w, ld), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().sbm_upsample_nearest2x(L.ptr(x), C.c_int64(x.stride(2)), L.ptr(out), C.c_int64(out.stride(2)),
                                           C.c_int32(b), C.c_int32(h), C.c_int32(w), C.c_int32(c), L.stream_ptr()),
            "sbm_upsample_nearest2x")
    return out


def colsum_per_sample(x: torch.Tensor, c: int, out: torch.Tensor) -> None:
    """out[b, :c] = sum over the pixels of sample b of x[b, :, :, :c] (out: a [B, >=c] fp32 view with row stride)."""
    b, h, w, _ = x.shape
    L.check(L.lib().sbm_colsum_per_sample(L.ptr(x), C.c_int32(_dt(x)), C.c_int64(x.stride(2)), C.c_int32(b),
                                          C.c_int32(h * w), C.c_int32(c), L.ptr(out), C.c_int64(out.stride(0)),
                                          L.stream_ptr()), "sbm_colsum_per_sample")


def upsample_nearest2x_bwd(dy: torch.Tensor, c: int) -> torch.Tensor:
    b, h2, w2, _ = dy.shape
    out = torch.empty((b, h2 // 2, w2 // 2, pad8(c)), dtype=torch.float32, device=dy.device)
    L.check(L.lib().sbm_upsample_nearest2x_bwd(L.ptr(dy), C.c_int64(dy.stride(2)), L.ptr(out), C.c_int64(out.stride(2)),
                                               None, C.c_int64(0), C.c_int32(b), C.c_int32(h2 // 2), C.c_int32(w2 // 2),
                                               C.c_int32(c), L.stream_ptr()), "sbm_upsample_nearest2x_bwd")
    return out


def scale_shift_bwd(n: torch.Tensor, dy: torch.Tensor, c: int, scale: torch.Tensor, shift: torch.Tensor, act: int,
                    dscale: torch.Tensor, dshift: torch.Tensor) -> torch.Tensor:
    """Backward of act(n * (1 + scale[b]) + shift[b]): returns dn (fp32, shaped like n); accumulates the per-sample
    dscale / dshift ([B, >=c] fp32 views sharing one row stride, zeroed by the caller)."""
    b, h, w, _ = n.shape
    assert scale.stride(0) == shift.stride(0) and dscale.stride(0) == dshift.stride(0)
    dn = torch.empty((b, h, w, pad8(c)), dtype=torch.float32, device=n.device)
    L.check(L.lib().sbm_scale_shift_bwd(L.ptr(n), C.c_int64(n.stride(2)), L.ptr(dy), C.c_int64(dy.stride(2)),
                                        L.ptr(scale), L.ptr(shift), C.c_int64(scale.stride(0)), C.c_int32(act),
                                        L.ptr(dn), C.c_int64(dn.stride(2)), L.ptr(dscale), L.ptr(dshift),
                                        C.c_int64(dscale.stride(0)), C.c_int32(b), C.c_int32(h * w), C.c_int32(c),
                                        L.stream_ptr()), "sbm_scale_shift_bwd")
    return dn
