"""GPU parity of the tcgen05 implicit-GEMM convolution (sbm_conv_igemm) against float64
torch convolutions on the SAME bf16-rounded operands, so the only difference is fp32
accumulation order (tolerance 2e-5 of the output scale).  Geometry cases follow the layer
shapes of the reference score nets (unet_model.py:30,33,107,110,208,272; unet_openai.py:185,207).
"""
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from score_based_multimodal_autoencoder_b200 import _lib as L, ops
    return L, ops


def _nhwc_bf16(x_nchw: torch.Tensor, ld: int) -> torch.Tensor:
    b, c, h, w = x_nchw.shape
    out = torch.full((b, h, w, ld), 7.0, dtype=torch.bfloat16, device=x_nchw.device)  # poison the padding
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def _gelu64(x):
    return 0.5 * x * (1.0 + torch.erf(x / 2.0 ** 0.5))


CASES = [
    # name,        kind, k, B,  H,  W,  cin, cout
    ("lin64",       "s1", 1, 128, 1, 1, 64, 64),
    ("lin256x128",  "s1", 1, 200, 1, 1, 256, 128),
    ("lin1024x512", "s1", 1, 64, 1, 1, 1024, 512),
    ("c1x1_16",     "s1", 1, 3, 16, 16, 256, 384),
    ("c3_16_64",    "s1", 3, 2, 16, 16, 64, 64),
    ("c3_16_tail",  "s1", 3, 3, 16, 16, 170, 512),
    ("c3_8_42",     "s1", 3, 5, 8, 8, 42, 128),
    ("c3_8_big",    "s1", 3, 4, 8, 8, 512, 256),
    ("c3_4",        "s1", 3, 9, 4, 4, 128, 256),
    ("c3_2",        "s1", 3, 33, 2, 2, 256, 128),
    ("c3_1",        "s1", 3, 70, 1, 1, 128, 256),
    ("c3_32x16",    "s1", 3, 2, 32, 16, 64, 96),
    ("c3_cout5",    "s1", 1, 4, 8, 8, 64, 5),
    ("down4_16",    "s2", 4, 3, 16, 16, 64, 64),
    ("down4_8",     "s2", 4, 5, 8, 8, 128, 128),
    ("down4_2",     "s2", 4, 40, 2, 2, 128, 128),
    ("down3_16",    "s2", 3, 3, 16, 16, 128, 128),
    ("up4_4",       "t",  4, 5, 4, 4, 128, 128),
    ("up4_1",       "t",  4, 50, 1, 1, 128, 128),
    ("up4_8",       "t",  4, 3, 8, 8, 256, 256),
    # large enough for the persistent CTA-pair (cta_group::2) kernel: >= 74 pairs of 128-row tiles
    ("pair_c3_16",   "s1", 3, 80, 16, 16, 64, 256),    # 160 m-tiles, BN=256
    ("pair_odd",     "s1", 3, 149, 8, 8, 72, 256),     # 75 m-tiles (odd): last pair half empty, cin tail
    ("pair_n512",    "s1", 1, 300, 8, 8, 128, 512),    # two n-tiles per pair row
    ("pair_bn128",   "s1", 3, 160, 16, 16, 64, 128),   # BN=128 pair variant
    ("pair_lin",     "s1", 1, 20000, 1, 1, 96, 256),   # linear layer, 157 m-tiles, ragged batch
    ("pair_down",    "s2", 4, 330, 16, 16, 64, 256),   # stride-2 view through the pair kernel
    ("pair_up",      "t",  4, 320, 4, 4, 64, 256),     # 4 output phases x 20 pairs... (80 pair tiles)
    ("pair_small2",  "s1", 3, 9000, 2, 2, 64, 128),    # 2x2 maps: 32 images per tile
    # BN=64 pair variant (cout <= 64: the PolyMNIST net's narrow layers at large batch)
    ("pair64_c3_8",  "s1", 3, 600, 8, 8, 64, 64),      # 300 m-tiles
    ("pair64_c42",   "s1", 3, 330, 8, 8, 42, 42),      # cin and cout tails (init_dim = 42)
    ("pair64_lin",   "s1", 1, 12000, 1, 1, 128, 64),
    ("pair64_down",  "s2", 4, 500, 8, 8, 64, 64),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_geometry(case):
    L, ops = _mods()
    name, kind, k, B, H, W, cin, cout = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()))
    x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    if kind == "t":
        w = (torch.randn(cin, cout, k, k, generator=g) / (cin * 4) ** 0.5).to(dev).to(torch.bfloat16).float()
    else:
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g).to(dev)
    xb = _nhwc_bf16(x, ops.pad8(cin) + 8)  # ld > pad8(cin): exercises the pixel stride
    if kind == "s1":
        ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)
        wpk = ops.pack_conv2d_weight(w)
        out = ops.conv_igemm(xb, wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, bias=bias)
    elif kind == "s2":
        ref = F.conv2d(x.double(), w.double(), bias.double(), stride=2, padding=1)
        wpk = ops.pack_conv2d_weight(w)
        out = ops.conv_igemm(xb, wpk, kind=L.CONV_S2, kh=k, kw=k, cin=cin, cout=cout, bias=bias)
    else:
        ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), stride=2, padding=1)
        wpk = ops.pack_convT2d_weight(w)
        out = ops.conv_igemm(xb, wpk, kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=cin, cout=cout, bias=bias)
    torch.cuda.synchronize()
    got = out[..., :cout].permute(0, 3, 1, 2).double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * scale + 1e-6, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


def test_conv_epilogue_options():
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    B, H, W, cin, cout = 6, 8, 8, 128, 192
    x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(B, H, W, cout, generator=g).to(dev)
    xb = _nhwc_bf16(x, cin)
    wpk = ops.pack_conv2d_weight(w)
    pre = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)

    # GELU + fp32 residual + statistics + bf16 side copy
    stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    out2 = torch.zeros(B, H, W, cout, dtype=torch.bfloat16, device=dev)
    out = ops.conv_igemm(xb, wpk, kind=L.CONV_S1, kh=3, kw=3, cin=cin, cout=cout, bias=bias, act=L.ACT_GELU,
                         residual=res, stats=stats, out2=out2)
    ref = _gelu64(pre) + res.double()
    torch.cuda.synchronize()
    assert (out.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()
    assert (out2.double() - ref).abs().max().item() <= 5e-3 * ref.abs().max().item()
    ref_s = torch.stack([ref.sum(dim=(1, 2, 3)), (ref * ref).sum(dim=(1, 2, 3))], dim=1)
    assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-3)

    # SiLU, bf16 output with statistics of the ROUNDED values
    stats.zero_()
    outb = ops.conv_igemm(xb, wpk, kind=L.CONV_S1, kh=3, kw=3, cin=cin, cout=cout, bias=bias, act=L.ACT_SILU,
                          out_dtype=torch.bfloat16, stats=stats)
    refs = pre * torch.sigmoid(pre)
    torch.cuda.synchronize()
    assert (outb.double() - refs).abs().max().item() <= 5e-3 * refs.abs().max().item()
    ob = outb.double()
    ref_s = torch.stack([ob.sum(dim=(1, 2, 3)), (ob * ob).sum(dim=(1, 2, 3))], dim=1)
    assert torch.allclose(stats, ref_s, rtol=1e-6, atol=1e-4)

    # NCHW fp32 output (final 1x1 projection to the latent channels)
    w1 = (torch.randn(5, cin, 1, 1, generator=g) / cin ** 0.5).to(dev).to(torch.bfloat16).float()
    b1 = torch.randn(5, generator=g).to(dev)
    o = ops.conv_igemm(xb, ops.pack_conv2d_weight(w1), kind=L.CONV_S1, kh=1, kw=1, cin=cin, cout=5, bias=b1, nchw=True)
    r = F.conv2d(x.double(), w1.double(), b1.double())
    torch.cuda.synchronize()
    assert o.shape == (B, 5, H, W)
    assert (o.double() - r).abs().max().item() <= 2e-5 * r.abs().max().item()


def test_conv_pair_kernel_epilogue_and_ab():
    """CTA-pair kernel with the full epilogue (GELU, residual, stats, bf16 copy) and A/B against the single-CTA kernel."""
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(21)
    B, H, W, cin, cout = 100, 16, 16, 64, 256
    x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(B, H, W, cout, generator=g).to(dev)
    xb = _nhwc_bf16(x, cin)
    wpk = ops.pack_conv2d_weight(w)
    outs = []
    for single in (1, 0):
        L.lib().sbm_conv_force_single_cta(single)
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        out2 = torch.zeros(B, H, W, cout, dtype=torch.bfloat16, device=dev)
        out = ops.conv_igemm(xb, wpk, kind=L.CONV_S1, kh=3, kw=3, cin=cin, cout=cout, bias=bias, act=L.ACT_GELU,
                             residual=res, stats=stats, out2=out2)
        torch.cuda.synchronize()
        outs.append((out, out2, stats))
    L.lib().sbm_conv_force_single_cta(0)
    pre = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    ref = _gelu64(pre) + res.double()
    for out, out2, stats in outs:
        assert (out.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()
        assert (out2.double() - ref).abs().max().item() <= 5e-3 * ref.abs().max().item()
        ref_s = torch.stack([ref.sum(dim=(1, 2, 3)), (ref * ref).sum(dim=(1, 2, 3))], dim=1)
        assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-3)
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-6)


def test_conv_small_stats_segments():
    """Per-sample statistics when several samples share one warp (OH*OW < 32)."""
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(9)
    for hw in (1, 2, 4):
        B, cin, cout = 37, 64, 64
        x = torch.randn(B, cin, hw, hw, generator=g).to(dev).to(torch.bfloat16).float()
        w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).to(torch.bfloat16).float()
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        out = ops.conv_igemm(_nhwc_bf16(x, cin), ops.pack_conv2d_weight(w), kind=L.CONV_S1, kh=3, kw=3, cin=cin,
                             cout=cout, stats=stats)
        torch.cuda.synchronize()
        o = out.double()
        ref_s = torch.stack([o.sum(dim=(1, 2, 3)), (o * o).sum(dim=(1, 2, 3))], dim=1)
        assert torch.allclose(stats, ref_s, rtol=1e-6, atol=1e-5), f"hw={hw}"
        ref = F.conv2d(x.double(), w.double(), padding=1).permute(0, 2, 3, 1)
        assert (o - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("variant", ["f32_res_out2", "bf16_gelu_preact", "bf16_out_f32_res", "bf16_res", "cat_views",
                                     "rowbias_silu"])
def test_conv_pair_staged_epilogue(variant):
    """TMA-staged epilogue of the CTA-pair kernel (swizzled shared-memory rows -> box stores, residual box loads)
    against float64 torch, and bit-for-bit against the per-thread-store epilogue (A/B switch)."""
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(variant.encode()))
    B, H, W, cin, cout = 150, 8, 8, 64, 256
    x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g).to(dev)
    xb = _nhwc_bf16(x, cin)
    wpk = ops.pack_conv2d_weight(w)
    pre = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    res32 = torch.randn(B, H, W, cout, generator=g).to(dev)
    rb = torch.randn(B, cout, generator=g).to(dev)
    results = []
    for direct in (1, 0):
        L.lib().sbm_conv_force_direct_epilogue(direct)
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        kw = dict(kind=L.CONV_S1, kh=3, kw=3, cin=cin, cout=cout, bias=bias, stats=stats)
        if variant == "f32_res_out2":
            out2 = torch.zeros(B, H, W, cout, dtype=torch.bfloat16, device=dev)
            out = ops.conv_igemm(xb, wpk, residual=res32, out2=out2, **kw)
            ref, ref2, tol = pre + res32.double(), pre + res32.double(), 3e-5
        elif variant == "bf16_gelu_preact":
            out2 = torch.zeros(B, H, W, cout, dtype=torch.bfloat16, device=dev)
            out = ops.conv_igemm(xb, wpk, act=L.ACT_GELU, out_dtype=torch.bfloat16, out2=out2, out2_preact=True, **kw)
            ref, ref2, tol = _gelu64(pre), pre, 5e-3
        elif variant == "bf16_out_f32_res":
            out2 = None
            out = ops.conv_igemm(xb, wpk, residual=res32, out_dtype=torch.bfloat16, **kw)
            ref, ref2, tol = pre + res32.double(), None, 5e-3
        elif variant == "bf16_res":
            out2 = None
            resb = res32.to(torch.bfloat16)
            out = ops.conv_igemm(xb, wpk, residual=resb, **kw)
            ref, ref2, tol = pre + resb.double(), None, 3e-5
        elif variant == "cat_views":
            cat_f = torch.full((B, H, W, 2 * cout), 3.0, dtype=torch.float32, device=dev)
            cat_b = torch.full((B, H, W, 2 * cout), 3.0, dtype=torch.bfloat16, device=dev)
            out = ops.conv_igemm(xb, wpk, residual=res32, out=cat_f[..., cout:], out2=cat_b[..., cout:], **kw)
            out2 = cat_b[..., cout:]
            ref, ref2, tol = pre + res32.double(), pre + res32.double(), 3e-5
            torch.cuda.synchronize()
            assert (cat_f[..., :cout] == 3.0).all() and (cat_b[..., :cout] == 3.0).all()  # neighbours untouched
        else:
            out2 = None
            out = ops.conv_igemm(xb, wpk, act=L.ACT_SILU, rowbias=rb, **kw)
            pr = pre + rb.double()[:, None, None, :]
            ref, ref2, tol = pr * torch.sigmoid(pr), None, 3e-5
        torch.cuda.synchronize()
        assert (out.double() - ref).abs().max().item() <= tol * ref.abs().max().item(), (variant, direct)
        if ref2 is not None:
            assert (out2.double() - ref2).abs().max().item() <= 5e-3 * ref2.abs().max().item(), (variant, direct)
        o = out.double()
        ref_s = torch.stack([o.sum(dim=(1, 2, 3)), (o * o).sum(dim=(1, 2, 3))], dim=1)
        assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-3), (variant, direct)
        results.append((out.clone(), None if out2 is None else out2.clone()))
    L.lib().sbm_conv_force_direct_epilogue(0)
    assert torch.equal(results[0][0], results[1][0])
    if results[0][1] is not None:
        assert torch.equal(results[0][1], results[1][1])


def test_conv_pair_staged_transposed_and_small_maps():
    """Staged epilogue box geometry: 4 output-parity phases of the transposed conv (q coordinate), 2x2 maps
    (8 images per warp box), 32-wide maps, ragged batch; out2 written into a channel-offset view."""
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(77)
    # transposed conv, output into the first half of a concat buffer
    B, H, W, c = 330, 4, 4, 256
    x = torch.randn(B, 64, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(64, c, 4, 4, generator=g) / 16.0).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(c, generator=g).to(dev)
    cat_f = torch.full((B, 2 * H, 2 * W, 2 * c), 3.0, dtype=torch.float32, device=dev)
    cat_b = torch.full((B, 2 * H, 2 * W, 2 * c), 3.0, dtype=torch.bfloat16, device=dev)
    ops.conv_igemm(_nhwc_bf16(x, 64), ops.pack_convT2d_weight(w), kind=L.CONVT_4X4_S2, kh=4, kw=4, cin=64, cout=c,
                   bias=bias, out=cat_f[..., :c], out2=cat_b[..., :c])
    ref = F.conv_transpose2d(x.double(), w.double(), bias.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert (cat_f[..., :c].double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert (cat_b[..., :c].double() - ref).abs().max().item() <= 5e-3 * ref.abs().max().item()
    assert (cat_f[..., c:] == 3.0).all() and (cat_b[..., c:] == 3.0).all()
    for (B, H, W, cin, cout) in [(9001, 2, 2, 64, 128), (41, 32, 32, 64, 256), (19999, 1, 1, 96, 256),
                                 (300, 16, 16, 64, 128)]:
        x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
        k = 3 if H > 1 else 1
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).to(torch.bfloat16).float()
        res = torch.randn(B, H, W, cout, generator=g).to(dev)
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        out = ops.conv_igemm(_nhwc_bf16(x, cin), ops.pack_conv2d_weight(w), kind=L.CONV_S1, kh=k, kw=k, cin=cin,
                             cout=cout, residual=res, stats=stats)
        ref = F.conv2d(x.double(), w.double(), padding=k // 2).permute(0, 2, 3, 1) + res.double()
        torch.cuda.synchronize()
        assert (out.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item(), (B, H, W)
        ref_s = torch.stack([ref.sum(dim=(1, 2, 3)), (ref * ref).sum(dim=(1, 2, 3))], dim=1)
        assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-3), (B, H, W)


@pytest.mark.parametrize("case", [
    # B,   H,  W, cin, cout, k
    (6,    8,  8, 96, 128, 3),     # single-CTA kernel, all 9 border classes
    (300,  8,  8, 64, 256, 3),     # CTA-pair kernel (staged epilogue)
    (700,  2,  2, 64, 128, 3),     # 2x2 maps: every pixel is a corner
    (200,  1,  1, 128, 64, 3),     # 1x1 maps: centre tap only
    (40,  16, 16, 72, 40, 1),      # 1x1 kernel, cout not a multiple of 16
    (90,  16, 16, 64, 256, 3),     # pair kernel on 16x16
], ids=lambda c: "x".join(map(str, c)))
def test_conv_folded_groupnorm(case):
    """GroupNorm(1,C) -> conv with the normalisation folded into the GEMM: gamma in the weights, mean / rstd / beta
    applied in the epilogue through per-border-class tables (unet_model.py:106-110 without the GroupNorm-apply pass)."""
    L, ops = _mods()
    B, H, W, cin, cout, k = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(B * 7 + H)
    x = (torch.randn(B, cin, H, W, generator=g) * 1.7 + 0.6).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    gamma = (1.0 + 0.3 * torch.randn(cin, generator=g)).to(dev)
    beta = (0.2 * torch.randn(cin, generator=g)).to(dev)
    res = torch.randn(B, H, W, cout, generator=g).to(dev)
    xd = x.double()
    ref = F.conv2d(F.group_norm(xd, 1, gamma.double(), beta.double(), eps=1e-5), w.double(), bias.double(),
                   padding=k // 2).permute(0, 2, 3, 1) + res.double()
    stats = torch.stack([xd.sum(dim=(1, 2, 3)), (xd * xd).sum(dim=(1, 2, 3))], dim=1).contiguous()
    wpk, tab = ops.fold_groupnorm_conv(w, bias, gamma, beta)
    st_out = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    out = ops.conv_igemm(_nhwc_bf16(x, ops.pad8(cin)), wpk, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout,
                         residual=res, gn_stats=stats, gn_tab=tab, stats=st_out)
    torch.cuda.synchronize()
    got = out[..., :cout].double()
    err = (got - ref).abs().max().item()
    assert err <= 6e-3 * ref.abs().max().item(), f"{case}: {err:.3e} vs {ref.abs().max().item():.3e}"
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 2.5e-3, rel
    ref_s = torch.stack([got.sum(dim=(1, 2, 3)), (got * got).sum(dim=(1, 2, 3))], dim=1)
    assert torch.allclose(st_out, ref_s, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("case", [
    # B,   H,  W, cin, cout, k, options
    (256, 16, 16, 64, 256, 3, "f32_res_out2"),      # one 256-sample block, all border classes of a 16x16 map
    (300,  8,  8, 72, 256, 3, "bf16_gelu_stats"),   # ragged batch: second sample block is mostly out of bounds
    (512,  4,  4, 128, 384, 3, "gnfold_res"),       # BN=128 pair, GroupNorm folded (per-pixel border class)
    (1024, 2,  2, 64, 512, 3, "rowbias_silu"),      # every pixel a corner: 4 of 9 taps
    (256,  8, 16, 64, 320, 3, "cat_views"),         # non-square map, cout tail, strided output views
    (256,  4,  4, 64, 256, 1, "f32_res_out2"),      # 1x1 kernel: nothing to skip, must stay on the standard tiling
], ids=lambda c: "x".join(map(str, c)))
def test_conv_pixel_major_tiling(case):
    """Pixel-major tiling (a tile = 128 samples at one output pixel; taps that only read zero padding there are skipped)
    against float64 torch and BIT-FOR-BIT against the standard tiling: the skipped products are exact zeros."""
    L, ops = _mods()
    B, H, W, cin, cout, k, opt = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(B * 131 + H * 7 + cout)
    x = (torch.randn(B, cin, H, W, generator=g) * 1.3 + 0.4).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(B, H, W, cout, generator=g).to(dev)
    rb = torch.randn(B, cout, generator=g).to(dev)
    xb = _nhwc_bf16(x, ops.pad8(cin))
    xd = x.double()
    outs = []
    for mode in (0, 1):
        L.lib().sbm_conv_pixel_major(mode)
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        kw = dict(kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout, stats=stats)
        out2 = None
        if opt == "gnfold_res":
            gamma = (1.0 + 0.3 * torch.randn(cin, generator=torch.Generator().manual_seed(1))).to(dev)
            beta = (0.2 * torch.randn(cin, generator=torch.Generator().manual_seed(2))).to(dev)
            wpk, tab = ops.fold_groupnorm_conv(w, bias, gamma, beta)
            gst = torch.stack([xd.sum(dim=(1, 2, 3)), (xd * xd).sum(dim=(1, 2, 3))], dim=1).contiguous()
            out = ops.conv_igemm(xb, wpk, residual=res, gn_stats=gst, gn_tab=tab, **kw)
            ref = F.conv2d(F.group_norm(xd, 1, gamma.double(), beta.double(), eps=1e-5), w.double(), bias.double(),
                           padding=k // 2).permute(0, 2, 3, 1) + res.double()
            tol = 6e-3
        else:
            wq = w.to(torch.bfloat16).float()
            wpk = ops.pack_conv2d_weight(wq)
            pre = F.conv2d(xd, wq.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
            if opt == "f32_res_out2":
                out2 = torch.zeros(B, H, W, ops.pad8(cout), dtype=torch.bfloat16, device=dev)
                out = ops.conv_igemm(xb, wpk, bias=bias, residual=res, out2=out2, **kw)
                ref, tol = pre + res.double(), 3e-5
            elif opt == "bf16_gelu_stats":
                out = ops.conv_igemm(xb, wpk, bias=bias, act=L.ACT_GELU, out_dtype=torch.bfloat16, **kw)
                ref, tol = _gelu64(pre), 5e-3
            elif opt == "rowbias_silu":
                out = ops.conv_igemm(xb, wpk, bias=bias, act=L.ACT_SILU, rowbias=rb, **kw)
                pr = pre + rb.double()[:, None, None, :]
                ref, tol = pr * torch.sigmoid(pr), 3e-5
            else:  # cat_views
                cat_f = torch.full((B, H, W, 2 * ops.pad8(cout)), 3.0, dtype=torch.float32, device=dev)
                cat_b = torch.full((B, H, W, 2 * ops.pad8(cout)), 3.0, dtype=torch.bfloat16, device=dev)
                o0 = ops.pad8(cout)
                out = ops.conv_igemm(xb, wpk, bias=bias, residual=res, out=cat_f[..., o0:], out2=cat_b[..., o0:], **kw)
                out2 = cat_b[..., o0:]
                ref, tol = pre + res.double(), 3e-5
                torch.cuda.synchronize()
                assert (cat_f[..., :o0] == 3.0).all() and (cat_b[..., :o0] == 3.0).all()
        torch.cuda.synchronize()
        variant = L.lib().sbm_conv_last_variant()
        assert bool(variant & (1 << 18)) == (mode == 1 and k > 1), (mode, hex(variant))
        got = out[..., :cout].double()
        assert (got - ref).abs().max().item() <= tol * ref.abs().max().item(), (case, mode)
        ref_s = torch.stack([got.sum(dim=(1, 2, 3)), (got * got).sum(dim=(1, 2, 3))], dim=1)
        assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-3), (case, mode)
        outs.append((out[..., :cout].clone(), None if out2 is None else out2[..., :cout].clone()))
    L.lib().sbm_conv_pixel_major(-1)
    assert torch.equal(outs[0][0], outs[1][0])
    if outs[0][1] is not None:
        assert torch.equal(outs[0][1], outs[1][1])
    # the A/B switch of the per-thread-store epilogue takes the same tiling
    if opt == "f32_res_out2" and k == 3:
        L.lib().sbm_conv_pixel_major(1)
        L.lib().sbm_conv_force_direct_epilogue(1)
        o3 = ops.conv_igemm(xb, wpk, bias=bias, residual=res, kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout)
        torch.cuda.synchronize()
        v = L.lib().sbm_conv_last_variant()
        L.lib().sbm_conv_force_direct_epilogue(0)
        L.lib().sbm_conv_pixel_major(-1)
        assert (v & (1 << 18)) and not (v & (1 << 17))
        assert torch.equal(o3[..., :cout], outs[0][0])


@pytest.mark.parametrize("opt", ["gn_gelu_bf16_stats", "gn_res_f32_out2_stats", "bias_f32", "bias_res_bf16out"])
@pytest.mark.parametrize("shape", [(4, 128, 1024, 512, 3), (2, 100, 512, 1024, 3), (8, 96, 640, 256, 3)])
def test_conv_split_k_matches_unsplit(opt, shape):
    """Sub-wave K-long layers (low-resolution levels at small batch) are cut along K across the SM pairs: partial sums
    go to per-slice slabs of an fp32 workspace, a second kernel sums them in slice order and applies the epilogue.  Same
    result as the unsplit launch up to fp32 summation order, every epilogue flag honoured; bit-identical run to run."""
    from score_based_multimodal_autoencoder_b200 import _lib as L, ops
    H, B, cin, cout, k = shape
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(H * 7 + cin)
    xb = (torch.randn(B, H, H, ops.pad8(cin), generator=g)).to(dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(B, H, H, ops.pad8(cout), generator=g).to(dev)
    kw = dict(kind=L.CONV_S1, kh=k, kw=k, cin=cin, cout=cout)
    gamma = (1 + 0.3 * torch.randn(cin, generator=g)).to(dev)
    beta = (0.2 * torch.randn(cin, generator=g)).to(dev)
    outs = []
    for split in (1, 1, 1, 0):   # the split launch three times: slabs summed in slice order -> bit-identical outputs
        L.lib().sbm_conv_splitk(split)
        stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
        out2 = None
        if opt.startswith("gn_"):
            wg, tab = ops.fold_groupnorm_conv(w, bias, gamma, beta)
            xs = xb[..., :cin].double()
            gst = torch.stack([xs.sum(dim=(1, 2, 3)), (xs * xs).sum(dim=(1, 2, 3))], dim=1).contiguous()
            gkw = dict(gn_stats=gst, gn_tab=tab, gn_eps=1e-5)
            if opt == "gn_gelu_bf16_stats":
                out = ops.conv_igemm(xb, wg, act=L.ACT_GELU, out_dtype=torch.bfloat16, stats=stats, **gkw, **kw)
            else:
                out2 = torch.empty(B, H, H, ops.pad8(cout), dtype=torch.bfloat16, device=dev)
                out = ops.conv_igemm(xb, wg, residual=res, stats=stats, out2=out2, **gkw, **kw)
        elif opt == "bias_f32":
            out = ops.conv_igemm(xb, ops.pack_conv2d_weight(w), bias=bias, **kw)
        else:
            out = ops.conv_igemm(xb, ops.pack_conv2d_weight(w), bias=bias, residual=res, out_dtype=torch.bfloat16, **kw)
        torch.cuda.synchronize()
        v = L.lib().sbm_conv_last_variant()
        assert bool(v & (1 << 20)) == bool(split), (split, hex(v))
        outs.append((out[..., :cout].float().clone(), None if out2 is None else out2[..., :cout].float().clone(), stats.clone()))
    L.lib().sbm_conv_splitk(1)
    a, a1, a2_, b_ = outs
    for again in (a1, a2_):
        assert torch.equal(a[0], again[0]) and (a[1] is None or torch.equal(a[1], again[1]))
        assert torch.allclose(a[2], again[2], rtol=1e-12, atol=1e-9)   # fp64 statistics: atomics, order-dependent last bits
    tol = 2e-2 if "bf16" in opt else 1e-4   # bf16 outputs: one rounding step of either result
    assert (a[0] - b_[0]).abs().max().item() <= tol * b_[0].abs().max().item()
    if a[1] is not None:
        assert (a[1] - b_[1]).abs().max().item() <= 2e-2 * b_[1].abs().max().item()
    assert torch.allclose(a[2], b_[2], rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("shape", [(8, 128, 512, 512, 4), (4, 96, 512, 512, 4), (2, 256, 512, 512, 4), (8, 64, 256, 256, 3)])
def test_conv_split_k_stride2_matches_unsplit(shape):
    """The stride-2 down-sampling convolutions (4x4 of unet_model.py:32-33, 3x3 of unet_openai.py:207) at small batch:
    16 (9) taps of K on a quarter of the pixels.  Split along K like the stride-1 layers; bias + bf16 copy epilogue (the
    flag set the nets use for them); result equal to the unsplit launch up to fp32 summation order."""
    from score_based_multimodal_autoencoder_b200 import _lib as L, ops
    H, B, cin, cout, k = shape
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(H * 11 + cout)
    xb = torch.randn(B, H, H, ops.pad8(cin), generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    wpk = ops.pack_conv2d_weight(w)
    outs = []
    for split in (1, 1, 1, 0):   # the split launch three times: the slabs are summed in slice order -> bit-identical
        L.lib().sbm_conv_splitk(split)
        o2 = torch.empty(B, H // 2, H // 2, ops.pad8(cout), dtype=torch.bfloat16, device=dev)
        out = ops.conv_igemm(xb, wpk, kind=L.CONV_S2, kh=k, kw=k, cin=cin, cout=cout, bias=bias, out2=o2)
        torch.cuda.synchronize()
        v = L.lib().sbm_conv_last_variant()
        assert bool(v & (1 << 20)) == bool(split), (split, hex(v))
        outs.append((out[..., :cout].clone(), o2[..., :cout].float().clone()))
    L.lib().sbm_conv_splitk(1)
    ref = torch.nn.functional.conv2d(xb[..., :cin].float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias,
                                     stride=2, padding=1).permute(0, 2, 3, 1)
    (a, a2), (a_again, a2_again), (a_third, a2_third), (b_, b2) = outs
    assert torch.equal(a, a_again) and torch.equal(a2, a2_again) and torch.equal(a, a_third) and torch.equal(a2, a2_third)
    scale = ref.abs().max().item()
    assert (a - b_).abs().max().item() <= 1e-4 * scale
    assert (a - ref).abs().max().item() <= 2e-3 * scale
    assert (a2 - b2).abs().max().item() <= 2e-2 * scale
