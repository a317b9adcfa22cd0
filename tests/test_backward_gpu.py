"""GPU parity of the backward kernels against float64 torch autograd on the same bf16-rounded operands."""
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from score_based_multimodal_autoencoder_b200 import _lib as L, ops
    return L, ops


def _nhwc_bf16(x_nchw, ld):
    b, c, h, w = x_nchw.shape
    out = torch.full((b, h, w, ld), 3.0, dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


WG_CASES = [
    # name, kind, k, B, H, W, cin, cout
    ("wg_lin", "s1", 1, 300, 1, 1, 96, 160),
    ("wg_c3_16", "s1", 3, 5, 16, 16, 64, 128),
    ("wg_c3_8_tail", "s1", 3, 9, 8, 8, 42, 200),
    ("wg_c3_4", "s1", 3, 33, 4, 4, 128, 256),
    ("wg_c3_2", "s1", 3, 70, 2, 2, 256, 128),
    ("wg_c3_1", "s1", 3, 130, 1, 1, 128, 256),
    ("wg_c1_16", "s1", 1, 3, 16, 16, 256, 384),
    ("wg_down_16", "s2", 4, 6, 16, 16, 64, 64),
    ("wg_down_2", "s2", 4, 50, 2, 2, 128, 128),
    ("wg_down3_8", "s2", 3, 7, 8, 8, 64, 128),
    ("wg_up_4", "t", 4, 9, 4, 4, 128, 64),
    ("wg_up_1", "t", 4, 70, 1, 1, 64, 128),
    ("wg_up_8", "t", 4, 4, 8, 8, 64, 64),
]


@pytest.mark.parametrize("case", WG_CASES, ids=[c[0] for c in WG_CASES])
def test_conv_wgrad(case):
    L, ops = _mods()
    name, kind, k, B, H, W, cin, cout = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()))
    x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).double()
    if kind == "t":
        w = torch.randn(cin, cout, k, k, generator=g).to(dev).double().requires_grad_(True)
        y = F.conv_transpose2d(x, w, stride=2, padding=1)
        knd = L.CONVT_4X4_S2
    elif kind == "s2":
        w = torch.randn(cout, cin, k, k, generator=g).to(dev).double().requires_grad_(True)
        y = F.conv2d(x, w, stride=2, padding=1)
        knd = L.CONV_S2
    else:
        w = torch.randn(cout, cin, k, k, generator=g).to(dev).double().requires_grad_(True)
        y = F.conv2d(x, w, padding=k // 2)
        knd = L.CONV_S1
    dy = torch.randn(y.shape, generator=g).to(dev).to(torch.bfloat16).double()
    (y * dy).sum().backward()
    xb = _nhwc_bf16(x.float(), ops.pad8(cin) + 8)
    dyb = _nhwc_bf16(dy.float(), ops.pad8(cout))
    dwpk = ops.conv_wgrad(xb, dyb, kind=knd, kh=k, kw=k, cin=cin, cout=cout)
    got = ops.unpack_convT2d_wgrad(dwpk, w) if kind == "t" else ops.unpack_conv2d_wgrad(dwpk, w)
    torch.cuda.synchronize()
    err = (got.double() - w.grad).abs().max().item()
    scale = w.grad.abs().max().item()
    assert err <= 3e-5 * scale + 1e-6, f"{name}: err {err:.3e} scale {scale:.3e}"
