"""TEST INFRASTRUCTURE: torch (CPU) stand-ins for the kernels the autoencoder plans call -- same packed bf16 operands,
same rounding points, fp32 accumulation -- so the HOST-side logic of h_vae_model_copy.py / h_vae_model.py (BatchNorm
folding, Linear-weight permutations, channel slicing, call order) is checked on CPU every round.  The kernels
themselves are checked on the GPU (tests/test_conv_igemm.py, tests/test_res_ae_gpu.py::test_*_resample_kernel)."""
import contextlib

import torch
import torch.nn.functional as F

from score_based_multimodal_autoencoder_b200 import _lib as L
from score_based_multimodal_autoencoder_b200 import h_vae_model as hm
from score_based_multimodal_autoencoder_b200 import h_vae_model_copy as hc
from score_based_multimodal_autoencoder_b200 import ops


def _pad8(n):
    return (n + 7) // 8 * 8


def pack_weight(w, taps, rows, cols, s_tap, s_row, s_col, out=None):
    src = torch.as_strided(w.contiguous().view(-1), (taps, rows, cols), (s_tap, s_row, s_col))
    o = torch.zeros(taps, rows, _pad8(cols), dtype=torch.bfloat16)
    o[..., :cols] = src.to(torch.bfloat16)
    return o


def conv_igemm(x, wpk, *, kind, kh, kw, cin, cout, bias=None, act=0, residual=None, out=None,
               out_dtype=torch.float32, nchw=False, **_unused):
    assert kind == L.CONV_S1 and act == 0 and x.dtype == torch.bfloat16
    xin = x[..., :cin].float().permute(0, 3, 1, 2)
    w = wpk[:, :cout, :cin].float().permute(1, 2, 0).reshape(cout, cin, kh, kw)
    y = F.conv2d(xin, w, None if bias is None else bias.float(), padding=(kh // 2, kw // 2))
    if residual is not None:
        y = y + residual[..., :cout].float().permute(0, 3, 1, 2)
    if nchw:
        return y.contiguous()
    o = torch.full((y.shape[0], y.shape[2], y.shape[3], _pad8(cout)), float("nan"), dtype=out_dtype)  # padding = garbage
    o[..., :cout] = y.permute(0, 2, 3, 1).to(out_dtype)
    return o


def stem_im2col(x, kh, kw):
    b, c, h, w = x.shape
    cols = F.unfold(x, (kh, kw), padding=(kh // 2, kw // 2)).transpose(1, 2).reshape(b, h, w, c * kh * kw)
    o = torch.zeros(b, h, w, _pad8(c * kh * kw), dtype=torch.bfloat16)
    o[..., :c * kh * kw] = cols.to(torch.bfloat16)
    return o


def nchw_to_nhwc(x, want_f32=False):
    b, c, h, w = x.shape
    o = torch.zeros(b, h, w, _pad8(c), dtype=torch.bfloat16)
    o[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return o, None


def lrelu_resample(x, c, slope, mode=0, rate=1, nchw=False, act=0):
    v = x[..., :c].float().permute(0, 3, 1, 2)
    v = F.gelu(v) if act == 1 else F.leaky_relu(v, slope)
    if mode == 1:
        v = F.avg_pool2d(v, rate)
    elif mode == 2:
        v = F.interpolate(v, scale_factor=rate, mode="nearest")
    elif mode == 3:
        v = F.interpolate(v, scale_factor=rate, mode="bilinear")
    if nchw:
        return v.contiguous()
    o = torch.zeros(v.shape[0], v.shape[2], v.shape[3], _pad8(c), dtype=torch.bfloat16)
    o[..., :c] = v.permute(0, 2, 3, 1).to(torch.bfloat16)
    return o


@contextlib.contextmanager
def emulated_kernels():
    """Swap the kernel wrappers for the torch stand-ins (and lift the CUDA-tensor check) inside the block."""
    saved = [(ops, n, getattr(ops, n)) for n in ("pack_weight", "conv_igemm", "stem_im2col", "nchw_to_nhwc")]
    saved += [(hc, "lrelu_resample", hc.lrelu_resample), (hm, "lrelu_resample", hm.lrelu_resample),
              (hc._ResBase, "_check", hc._ResBase._check), (hm._AttrBase, "_check", hm._AttrBase._check)]
    ops.pack_weight, ops.conv_igemm, ops.stem_im2col, ops.nchw_to_nhwc = pack_weight, conv_igemm, stem_im2col, nchw_to_nhwc
    hc.lrelu_resample = hm.lrelu_resample = lrelu_resample
    hc._ResBase._check = lambda self, t: None
    hm._AttrBase._check = lambda self, t: None
    try:
        yield
    finally:
        for obj, name, val in saved:
            setattr(obj, name, val)
