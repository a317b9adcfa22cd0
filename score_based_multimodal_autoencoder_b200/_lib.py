"""ctypes binding of libsbmae_b200.so (the C ABI declared in include/sbmae_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsbmae_b200.so")

_lib = None


class SbmError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SbmError(
                f"{LIB_PATH} is missing: build it with `python -m score_based_multimodal_autoencoder_b200.build` "
                "(there is no CPU / eager fallback for the B200 path)"
            )
        _lib = C.CDLL(LIB_PATH)
        _lib.sbm_last_error.restype = C.c_char_p
        _lib.sbm_launch_count.restype = C.c_ulonglong
        _lib.sbm_conv_splitk_ws_elems.restype = C.c_int64
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise SbmError(f"{what} failed (rc={rc}): {lib().sbm_last_error().decode()}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise SbmError("libsbmae_b200 operates on CUDA tensors only (no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(lib().sbm_launch_count())


# ------------------------------------------------------------------ enums (mirror sbmae_b200.h)
CONV_S1, CONV_S2, CONVT_4X4_S2 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
F32, BF16 = 0, 1


class ConvArgs(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("wpk", C.c_void_p), ("cin_pad", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out_dtype", C.c_int32), ("out_nchw", C.c_int32),
        ("res_dtype", C.c_int32), ("out2_preact", C.c_int32),
        ("stats", C.c_void_p),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("rowbias", C.c_void_p), ("ld_rowbias", C.c_int64),
        ("gn_stats", C.c_void_p), ("gn_tab", C.c_void_p), ("gn_count", C.c_float), ("gn_eps", C.c_float),
        ("splitk_ws", C.c_void_p), ("ld_ws", C.c_int64), ("ws_elems", C.c_int64),
    ]

SDE_VP, SDE_SUBVP, SDE_VE = 0, 1, 2


class LatentShape(C.Structure):
    _fields_ = [("batch", C.c_int32), ("mods", C.c_int32), ("dd", C.c_int32)]


class SdeC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("b0", C.c_float), ("b1", C.c_float), ("N", C.c_int32), ("T", C.c_float)]


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("draw", C.c_uint64), ("sample_offset", C.c_uint64), ("draw_dev", C.c_void_p)]


class Impute(C.Structure):
    _fields_ = [("z_obs", C.c_void_p), ("obs_mask", C.c_uint32), ("noise_obs", C.c_int32), ("t_next", C.c_float),
                ("t_next_dev", C.c_void_p)]


class WgradArgs(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("dy", C.c_void_p), ("lddy", C.c_int64),
        ("dwpk", C.c_void_p), ("cin_pad", C.c_int32), ("reserved", C.c_int32),
    ]


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64)]


class EmaTensor(C.Structure):
    _fields_ = [("ema", C.c_void_p), ("src", C.c_void_p), ("n", C.c_int64)]


class PackDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("taps", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("cols_pad", C.c_int32), ("s_tap", C.c_int64), ("s_row", C.c_int64), ("s_col", C.c_int64),
                ("tiles_c", C.c_int32), ("first_block", C.c_int32), ("taps_magic", C.c_uint32), ("reserved", C.c_int32)]


# ------------------------------------------------------------------ NVTX ranges (opt-in: SBM_NVTX=1)
import contextlib as _contextlib

_NVTX = os.environ.get("SBM_NVTX", "0") == "1"


@_contextlib.contextmanager
def nvtx(name: str):
    """Named range around a phase of the path (sampler step, score-net forward, training step) for Nsight timelines;
    a no-op unless SBM_NVTX=1 (ranges cost a host call each)."""
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
