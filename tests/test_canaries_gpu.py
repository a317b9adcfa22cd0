"""Out-of-bounds WRITE canaries for the kernels added or rewritten in round 2 (compute-sanitizer is disabled on this GPU
pool, profiles/r2_sanitizer_unavailable.txt): the output of each kernel is a view in the middle of a larger allocation
filled with a sentinel; after the call the guard bands on both sides (and row padding the kernel must not touch) still
hold the sentinel bit pattern.  Sizes are deliberately ragged (tails of tiles, blocks and warps)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096       # elements on each side


def _guarded(shape, dtype, sentinel):
    n = 1
    for d in shape:
        n *= d
    buf = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device="cuda")
    return buf, buf[GUARD:GUARD + n].view(shape)


def _intact(buf, sentinel):
    ref = torch.full((GUARD,), sentinel, dtype=buf.dtype, device="cuda")
    return bool(torch.equal(buf[:GUARD].view(torch.uint8), ref.view(torch.uint8)) and
                torch.equal(buf[-GUARD:].view(torch.uint8), ref.view(torch.uint8)))


def _mods():
    from score_based_multimodal_autoencoder_b200 import _lib as L, ops, sde_helper2 as sh
    return L, ops, sh


@pytest.mark.parametrize("shape", [(7, 5, 8, 8), (33, 3, 16, 16), (1, 1, 2, 2), (130, 5, 4, 4)])
def test_sampler_step_kernels_stay_inside_their_outputs(shape):
    L, ops, sh = _mods()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g).cuda()
    s = torch.randn(*shape, generator=g).cuda()
    t = (torch.rand(shape[0], generator=g) * 0.9 + 0.05).cuda()
    sent = -123.25
    for kind, a, b in (("vp", 0.1, 20.0), ("ve", 0.01, 5.0), ("subvp", 0.1, 20.0)):
        sde = {"vp": sh.VPSDE, "ve": sh.VESDE, "subvp": sh.subVPSDE}[kind](a, b, 50)
        for predictor in ("euler", "reverse_diffusion"):
            buf, out = _guarded(shape, torch.float32, sent)
            rng = sh._RngState()
            xn, xm = sh._predictor_kernel(sde, x, s, t, rng=rng.next(), out=out, predictor=predictor)
            torch.cuda.synchronize()
            assert _intact(buf, sent) and torch.isfinite(out).all()
        buf, out = _guarded(shape, torch.float32, sent)
        acc_buf, acc = _guarded((3,), torch.float64, -7.0)
        acc.zero_()
        rng = sh._RngState()
        r = rng.next()
        side = sh._fork_noise_norm(x, r, acc)
        torch.cuda.current_stream().wait_stream(side)
        sh._corrector_kernels(sde, x, s, t, 0.16, rng=r, acc=acc, out=out, noise_norm_done=True)
        torch.cuda.synchronize()
        assert _intact(buf, sent) and _intact(acc_buf, -7.0) and torch.isfinite(out).all()
        assert float(acc.abs().sum()) == 0.0          # the update kernel re-zeroed the accumulator


@pytest.mark.parametrize("B,n_side,heads", [(3, 16, 4), (5, 8, 4), (9, 4, 4), (70, 2, 4), (33, 1, 4), (2, 8, 1)])
def test_linear_attention_kernels_stay_inside_their_outputs(B, n_side, heads):
    """mma kernel (n = 64 / 256), warp-per-(head, sample) kernel (n <= 16): output rows of `ld` > heads*32 elements whose
    padding columns must keep the sentinel."""
    L, ops, sh = _mods()
    n = n_side * n_side
    hid = heads * 32
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(B, n_side, n_side, 3 * hid, generator=g).cuda()
    ld = hid + 8
    sent = 3.0
    buf, out = _guarded((B, n_side, n_side, ld), torch.bfloat16, sent)
    L.check(L.lib().sbm_linear_attn_fwd(L.ptr(qkv), C.c_int64(qkv.stride(2)), L.ptr(out), C.c_int64(ld), C.c_int32(B),
                                        C.c_int32(n), C.c_int32(heads), C.c_float(32 ** -0.5), L.stream_ptr()),
            "sbm_linear_attn_fwd")
    torch.cuda.synchronize()
    assert _intact(buf, sent)
    assert (out[..., hid:] == sent).all() and torch.isfinite(out[..., :hid].float()).all()
    ref = ops.linear_attn(qkv, heads, 32 ** -0.5)
    assert torch.equal(out[..., :hid], ref[..., :hid])


@pytest.mark.parametrize("B,C_,H", [(5, 3, 16), (64, 5, 8), (3, 1, 4)])
def test_stem_im2col_stays_inside_its_output(B, C_, H):
    L, ops, sh = _mods()
    x = torch.randn(B, C_, H, H, generator=torch.Generator().manual_seed(3)).cuda()
    ldk = ops.pad8(C_ * 49)
    sent = 5.0
    buf, out = _guarded((B, H, H, ldk), torch.bfloat16, sent)
    L.check(L.lib().sbm_stem_im2col(L.ptr(x), L.ptr(out), C.c_int32(B), C.c_int32(C_), C.c_int32(H), C.c_int32(H),
                                    C.c_int32(7), C.c_int32(7), C.c_int32(ldk), L.stream_ptr()), "sbm_stem_im2col")
    torch.cuda.synchronize()
    assert _intact(buf, sent)
    ref = torch.nn.functional.unfold(x, 7, padding=3).transpose(1, 2).reshape(B, H, H, C_ * 49)
    assert torch.equal(out[..., :C_ * 49].float(), ref.to(torch.bfloat16).float())
    assert (out[..., C_ * 49:] == 0).all()


def test_guidance_and_langevin_kernels_stay_inside_their_outputs():
    L, ops, sh = _mods()
    from score_based_multimodal_autoencoder_b200 import eval_samplers as es
    g = torch.Generator().manual_seed(4)
    B, M, D = 37, 3, 16
    x = torch.randn(B, M, D, D, generator=g).cuda()
    ls = L.LatentShape(B, M, D * D)
    ld = ops.pad8(2 * D * D) + 8
    buf, rows = _guarded((B, ld), torch.bfloat16, 9.0)
    L.check(L.lib().sbm_guidance_gather(C.byref(ls), L.ptr(x), C.c_int32(0), C.c_int32(2), L.ptr(rows), C.c_int32(ld),
                                        L.stream_ptr()), "sbm_guidance_gather")
    torch.cuda.synchronize()
    assert _intact(buf, 9.0)
    want = torch.cat((x[:, 0].reshape(B, -1), x[:, 2].reshape(B, -1)), 1).to(torch.bfloat16)
    assert torch.equal(rows[:, :2 * D * D], want) and (rows[:, 2 * D * D:] == 0).all()
    sbuf, score = _guarded((B, M, D, D), torch.float32, -2.0)
    score.copy_(torch.randn(B, M, D, D, generator=g))
    before = score.clone()
    grad = torch.randn(B, 2 * D * D, generator=g).cuda()
    L.check(L.lib().sbm_guidance_apply(C.byref(ls), L.ptr(score), L.ptr(grad), C.c_int64(2 * D * D), C.c_int32(-1),
                                       C.c_int32(1), C.c_float(0.5), L.stream_ptr()), "sbm_guidance_apply")
    torch.cuda.synchronize()
    assert _intact(sbuf, -2.0)
    assert torch.equal(score[:, 0], before[:, 0]) and torch.equal(score[:, 2], before[:, 2])
    assert torch.allclose(score[:, 1], before[:, 1] - 0.5 * grad[:, D * D:].view(B, D, D))
    obuf, out = _guarded((B, M, D, D), torch.float32, -4.0)
    nz = torch.randn(B, M, D, D, generator=g).cuda()
    es._axpy_step(x, before, [0.1, 0.2, 0.3], [0.01, 0.02, 0.03], 0b001, noise=nz, out=out)
    torch.cuda.synchronize()
    assert _intact(obuf, -4.0) and torch.equal(out[:, 0], x[:, 0])
    assert torch.allclose(out[:, 2], x[:, 2] + 0.3 * before[:, 2] + 0.03 * nz[:, 2], rtol=1e-6, atol=1e-6)
