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
    return out


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


def pack_linear_weight(w: torch.Tensor, out=None) -> torch.Tensor:
    """nn.Linear weight [O, I] -> [1, O, pad8(I)] bf16."""
    w = w.detach().contiguous()
    o, i = w.shape
    return pack_weight(w, 1, o, i, 0, i, 1, out)


def conv_igemm(x: torch.Tensor, wpk: torch.Tensor, *, kind: int, kh: int, kw: int, cin: int, cout: int,
               bias: torch.Tensor | None = None, act: int = L.ACT_NONE, residual: torch.Tensor | None = None,
               out: torch.Tensor | None = None, out_dtype: torch.dtype = torch.float32, nchw: bool = False,
               stats: torch.Tensor | None = None, out2: torch.Tensor | None = None) -> torch.Tensor:
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
    L.check(L.lib().sbm_conv_igemm(C.byref(a), L.stream_ptr()), "sbm_conv_igemm")
    return out
