"""GPU parity of the backward kernels against float64 torch autograd on the same bf16-rounded operands."""
import zlib

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2

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


def _nhwc(x_nchw, dtype=torch.float32):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dtype)


@pytest.mark.parametrize("groups,in_act,C,H", [(1, 0, 42, 8), (1, 1, 64, 4), (32, 2, 64, 8), (1, 0, 256, 1)])
def test_groupnorm_backward(groups, in_act, C, H):
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(C + H)
    B = 5
    pre = torch.randn(B, C, H, H, generator=g).to(dev).double().requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev).double().requires_grad_(True)
    beta = torch.randn(C, generator=g).to(dev).double().requires_grad_(True)
    x = pre if in_act == 0 else (F.gelu(pre) if in_act == 1 else F.silu(pre))
    y = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    dy = torch.randn(y.shape, generator=g).to(dev).double()
    (y * dy).sum().backward()
    ld = ops.pad8(C)
    pre_n = torch.zeros(B, H, H, ld, device=dev)
    pre_n[..., :C] = _nhwc(pre.detach().float())
    dy_n = torch.zeros(B, H, H, ld, device=dev)
    dy_n[..., :C] = _nhwc(dy.float())
    xv = x.detach().float()
    xg = xv.reshape(B, groups, -1).double()
    stats = torch.stack([xg.sum(-1), (xg * xg).sum(-1)], dim=-1).reshape(B, groups, 2).contiguous()
    dx, dxb, dg, db = ops.groupnorm_bwd(pre_n, dy_n, C, stats, gamma.detach().float(), groups=groups, in_act=in_act,
                                        want_f32=True, want_bf16=True)
    torch.cuda.synchronize()
    ref = _nhwc(pre.grad.float())
    assert (dx[..., :C] - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-5
    assert (dxb[..., :C].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-5
    assert torch.allclose(dg, gamma.grad.float(), rtol=2e-4, atol=2e-4)
    assert torch.allclose(db, beta.grad.float(), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("H,C,B", [(16, 170, 3), (16, 40, 2), (8, 64, 7), (8, 512, 9), (4, 33, 21), (2, 64, 70), (1, 70, 300)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_dwconv7_forward(H, C, B, out_dtype):
    """Depthwise 7x7 + bias + per-sample time condition + GroupNorm statistics (unet_model.py:103-105) against float64
    torch: ragged batches (odd sample counts per shared-memory step), odd channel counts, both output types."""
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(H * 1000 + C + B)
    x = torch.randn(B, C, H, H, generator=g).to(dev)
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev)
    bias = torch.randn(C, generator=g).to(dev)
    cond = torch.randn(B, C, generator=g).to(dev)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=3, groups=C) + cond.double()[:, :, None, None]
    ld = ops.pad8(C)
    xn = torch.full((B, H, H, ld), 3.0, device=dev)
    xn[..., :C] = _nhwc(x)
    condp = torch.zeros(B, ld, device=dev)
    condp[:, :C] = cond
    stats = torch.zeros(B, 1, 2, dtype=torch.float64, device=dev)
    out = ops.dwconv7(xn, C, w.contiguous(), bias, condp, ld, stats, out_dtype=out_dtype)
    torch.cuda.synchronize()
    got = out[..., :C].double()
    refn = _nhwc(ref)
    tol = 1e-5 if out_dtype == torch.float32 else 5e-3
    assert (got - refn).abs().max().item() <= tol * refn.abs().max().item()
    ref_s = torch.stack([got.sum(dim=(1, 2, 3)), (got * got).sum(dim=(1, 2, 3))], dim=-1).view(B, 1, 2)
    assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("H,C", [(16, 40), (8, 64), (4, 33), (2, 64), (1, 70)])
def test_dwconv7_backward(H, C):
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(H * 100 + C)
    B = 6
    x = torch.randn(B, C, H, H, generator=g).to(dev).double().requires_grad_(True)
    w = (torch.randn(C, 1, 7, 7, generator=g) / 7).to(dev).double().requires_grad_(True)
    bias = torch.randn(C, generator=g).to(dev).double().requires_grad_(True)
    cond = torch.randn(B, C, generator=g).to(dev).double().requires_grad_(True)
    y = F.conv2d(x, w, bias, padding=3, groups=C) + cond[:, :, None, None]
    dy = torch.randn(y.shape, generator=g).to(dev).double()
    (y * dy).sum().backward()
    ld = ops.pad8(C)
    xn = torch.zeros(B, H, H, ld, device=dev); xn[..., :C] = _nhwc(x.detach().float())
    dyn = torch.zeros(B, H, H, ld, device=dev); dyn[..., :C] = _nhwc(dy.float())
    addend = torch.randn(B, H, H, ld, generator=g).to(dev)
    dx = ops.dwconv7_bwd_input(dyn, C, w.detach().float().contiguous(), addend=addend)
    dcond = torch.zeros(B, 1, 1, ld + 8, device=dev)
    dw, db = ops.dwconv7_wgrad(xn, dyn, C, dcond=dcond[:, :, :, 8:], ldc=dcond.stride(2))
    torch.cuda.synchronize()
    ref = _nhwc(x.grad.float()) + addend[..., :C]
    assert (dx[..., :C] - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert torch.allclose(dw, w.grad.float(), rtol=1e-4, atol=1e-3)
    assert torch.allclose(db, bias.grad.float(), rtol=1e-4, atol=1e-3)
    assert torch.allclose(dcond[:, 0, 0, 8:8 + C], cond.grad.float(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("n_side", [1, 2, 4, 8, 16])
def test_linear_attention_backward(n_side):
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(n_side)
    B, heads, d = 3, 4, 32
    n = n_side * n_side
    qkv = torch.randn(B, n, 3 * heads * d, generator=g).to(dev).double().requires_grad_(True)
    q, k, v = (t.reshape(B, n, heads, d).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=-1))  # b h d n
    qs = q.softmax(dim=-2) * d ** -0.5
    ks = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", ks, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, qs)  # b h e n
    out_n = out.permute(0, 3, 1, 2).reshape(B, n, heads * d)
    do = torch.randn(out_n.shape, generator=g).to(dev).double()
    (out_n * do).sum().backward()
    qkv_t = qkv.detach().float().reshape(B, n_side, n_side, -1).contiguous()
    fwd = ops.linear_attn(qkv_t, heads, d ** -0.5)
    dqkv = ops.linear_attn_bwd(qkv_t, do.float().reshape(B, n_side, n_side, -1).contiguous(), heads, d ** -0.5)
    torch.cuda.synchronize()
    assert (fwd.float().reshape(B, n, -1) - out_n.detach().float()).abs().max().item() <= 1e-2 * out_n.abs().max().item()
    # bf16 output of an fp32 / tf32 computation: the error is rounding noise, not a layout slip
    assert rel_l2(fwd.float().reshape(B, n, -1), out_n.detach().float()) < 4e-3
    ref = qkv.grad.float()
    got = dqkv.float().reshape(B, n, -1)
    assert (got - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-6
    if n in (64, 256):
        # bf16 q | k | v (what the to_qkv GEMM hands over at 8x8 / 16x16): same kernel, the rounding of its INPUT is the
        # only difference -- against the reference evaluated on the same rounded input the error is the fp32 kernel's
        qb = qkv_t.to(torch.bfloat16)
        fwd_b = ops.linear_attn(qb, heads, d ** -0.5)
        fwd_f = ops.linear_attn(qb.float(), heads, d ** -0.5)
        torch.cuda.synchronize()
        assert torch.equal(fwd_b, fwd_f)
        assert rel_l2(fwd_b.float().reshape(B, n, -1), out_n.detach().float()) < 1.2e-2


@pytest.mark.parametrize("n_side,heads,dh,layout", [(1, 4, 32, "unet"), (4, 4, 32, "unet"), (4, 2, 48, "openai"),
                                                    (8, 1, 64, "openai")])
def test_softmax_attention_backward(n_side, heads, dh, layout):
    L, ops = _mods()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(n_side * 7 + dh)
    B = 3
    n = n_side * n_side
    hid = heads * dh
    qkv = torch.randn(B, n, 3 * hid, generator=g).to(dev).double().requires_grad_(True)
    if layout == "unet":
        q, k, v = (t.reshape(B, n, heads, dh) for t in qkv.chunk(3, dim=-1))
        offs = (0, hid, 2 * hid, dh)
        scale = dh ** -0.5
    else:
        t = qkv.reshape(B, n, heads, 3, dh)
        q, k, v = t[:, :, :, 0], t[:, :, :, 1], t[:, :, :, 2]
        offs = (0, dh, 2 * dh, 3 * dh)
        scale = dh ** -0.5
    sim = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bjhd->bihd", attn, v).reshape(B, n, hid)
    do = torch.randn(out.shape, generator=g).to(dev).double()
    (out * do).sum().backward()
    qkv_t = qkv.detach().float().reshape(B, n_side, n_side, -1).contiguous()
    fwd = ops.softmax_attn(qkv_t, heads, dh, offs[0], offs[1], offs[2], offs[3], scale)
    dqkv = ops.softmax_attn_bwd(qkv_t, do.float().reshape(B, n_side, n_side, -1).contiguous(), heads, dh, offs[0],
                                offs[1], offs[2], offs[3], scale, 3 * hid)
    torch.cuda.synchronize()
    assert (fwd.float().reshape(B, n, -1)[..., :hid] - out.detach().float()).abs().max().item() <= 1e-2 * out.abs().max().item()
    ref = qkv.grad.float()
    got = dqkv.float().reshape(B, n, -1)[..., :3 * hid]
    assert (got - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-6


def test_dsm_training_step_gradients_vs_reference_golden():
    """loss_fn(...).backward() through the B200 Unet against gradients of the real reference (tests/golden/dsm_loss.pt)."""
    from oracle.det_weights import fill_state_dict
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    from tests.util import golden, rel_l2
    g = golden("dsm_loss.pt")
    net = golden("unet_poly.pt")
    m = Unet(**net["kwargs"])
    m.load_state_dict(fill_state_dict(net["shapes"]))
    m = m.cuda().train()
    cls = {"vp": sh.VPSDE, "subvp": sh.subVPSDE, "ve": sh.VESDE}
    for c in g["cases"][:3]:
        sde = cls[c["kind"]](c["a"], c["b"], c["N"])
        m.zero_grad()
        loss = sh.loss_fn(g["batch"].cuda(), m, sde, reduce_mean=c["reduce_mean"],
                          likelihood_weighting=c["likelihood_weighting"], u=g["u"].cuda(), z=g["z"].cuda())
        loss.backward()
        torch.cuda.synchronize()
        assert abs(loss.item() - c["loss"].item()) <= 2e-2 * abs(c["loss"].item())
        params = dict(m.named_parameters())
        worst = 0.0
        for k, ref in c["grads"].items():
            got = params[k].grad
            assert got is not None, k
            head = got.flatten()[:512].float().cpu()
            rel = ((head - ref["head"]).norm() / (ref["head"].norm() + 1e-12)).item()
            nrel = abs(got.norm().item() - ref["norm"].item()) / (ref["norm"].item() + 1e-12)
            print(f"{c['kind']} rm={c['reduce_mean']} lw={c['likelihood_weighting']} {k}: head rel {rel:.3e}, norm rel {nrel:.3e}")
            worst = max(worst, rel)
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None)).item()
        print(f"  total grad norm {gn:.5e} vs reference {c['grad_norm'].item():.5e}")
        assert all(p.grad is not None for p in m.parameters())
        assert abs(gn - c["grad_norm"].item()) <= 3e-2 * c["grad_norm"].item()
        assert worst <= 6e-2, worst


def test_fused_adam_matches_torch_adam():
    from score_based_multimodal_autoencoder_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(300, 17), (5,), (70000,), (64, 3, 3, 3)]
    p_ref = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    p_new = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam(p_ref, lr=5e-4)
    o_new = FusedAdam(p_new, lr=5e-4)
    for it in range(5):
        for a, b in zip(p_ref, p_new):
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad = gr.clone()
            b.grad = gr.clone()
        o_ref.step()
        o_new.step()
    for a, b in zip(p_ref, p_new):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_fused_adam_capturable_eager_loop_state_dict_and_resume():
    """(1) FusedAdam(capturable=True) in a plain eager loop advances its device step counter itself (bias
    correction not frozen at step 1); (2) state_dict() reports the true step; (3) load_state_dict of a torch.optim.Adam
    checkpoint works (float-tensor steps) and the kernel then writes the LOADED moment tensors."""
    from score_based_multimodal_autoencoder_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(1)
    shapes = [(300, 17), (70000,)]
    p_ref = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    p_cap = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam(p_ref, lr=5e-4)
    o_cap = FusedAdam(p_cap, lr=5e-4, capturable=True)
    grads = [[torch.randn(s, generator=g).cuda() for s in shapes] for _ in range(9)]
    for it in range(6):
        for a, b, gr in zip(p_ref, p_cap, grads[it]):
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step()
        o_cap.step()
    for a, b in zip(p_ref, p_cap):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    sd = o_cap.state_dict()
    assert all(int(st["step"]) == 6 for st in sd["state"].values())
    # resume a fresh FusedAdam from torch.optim.Adam's checkpoint and continue in lock-step
    p_new = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_new = FusedAdam(p_new, lr=5e-4)
    for b, gr in zip(p_new, grads[0]):   # one throw-away step so that tables / moments exist before the load
        b.grad = gr.clone()
    o_new.step()
    with torch.no_grad():
        for b, a in zip(p_new, p_ref):
            b.copy_(a)
    import copy
    o_new.load_state_dict(copy.deepcopy(o_ref.state_dict()))   # (load_state_dict keeps same-device tensors by reference)
    for it in range(6, 9):
        for a, b, gr in zip(p_ref, p_new, grads[it]):
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step()
        o_new.step()
    for a, b in zip(p_ref, p_new):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    for a, b in zip(p_ref, p_new):
        assert torch.allclose(o_ref.state[a]["exp_avg_sq"], o_new.state[b]["exp_avg_sq"], rtol=1e-4, atol=1e-9)


# DSM-loss curve of the UNMODIFIED reference (unet_model.Unet(dim=32, channels=5, dim_mults=(1,2)) under torch.manual_seed(0),
# torch.optim.Adam(lr=5e-4), fixed batch / u / z from torch.Generator().manual_seed(3), fp32 CPU), generated in the build
# container with the reference modules imported from /root/reference (same recipe as oracle/gen_golden.py).
REF_TRAIN_CURVE = [1.0598, 1.1002, 1.0527, 0.9973, 0.9806, 0.9656, 0.9449, 0.9243, 0.9035, 0.8833, 0.8615, 0.8408]


def test_dsm_training_tracks_reference_loss_curve():
    """End-to-end training: loss_fn -> hand-written backward -> FusedAdam (and torch.optim.Adam) for 12 steps reproduces
    the reference's own loss curve.  This also pins that optimizer updates reach the packed bf16 GEMM operands (the
    weight cache is keyed by the parameter version, which FusedAdam bumps after writing through raw pointers)."""
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.optim import FusedAdam
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    g = torch.Generator().manual_seed(3)
    x = torch.randn(32, 5, 8, 8, generator=g).cuda()
    u = torch.rand(32, generator=g).cuda()
    z = torch.randn(32, 5, 8, 8, generator=g).cuda()
    sde = sh.VPSDE(1.0, 5.0, 100)
    for opt_cls in (FusedAdam, torch.optim.Adam):
        torch.manual_seed(0)
        m = Unet(dim=32, channels=5, dim_mults=(1, 2)).cuda().train()
        opt = opt_cls(m.parameters(), lr=5e-4)
        losses = []
        for it in range(12):
            loss = sh.loss_fn(x, m, sde, likelihood_weighting=False, u=u, z=z)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        print(opt_cls.__name__, [round(v, 4) for v in losses])
        assert losses[-1] < 0.85 * losses[0]
        for got, ref in zip(losses, REF_TRAIN_CURVE):
            assert abs(got - ref) <= 2e-2 * ref, (opt_cls.__name__, losses)


def test_graphed_train_step_equals_eager_steps():
    """The CUDA-graph replay of the training step (loss_fn + backward + FusedAdam, device-side Philox draw id and Adam
    step count) produces the same losses and parameters as the eager loop with the same seeds."""
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.optim import FusedAdam, GraphedTrainStep
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    g = torch.Generator().manual_seed(4)
    batches = [torch.randn(16, 5, 8, 8, generator=g).cuda() for _ in range(8)]
    sde = sh.VPSDE(1.0, 5.0, 100)
    kw = dict(dim=32, channels=5, dim_mults=(1, 2))

    torch.manual_seed(0)
    m_e = Unet(**kw).cuda().train()
    opt = FusedAdam(m_e.parameters(), lr=5e-4)
    sh.manual_seed(99)
    eager = []
    for b in [batches[0], batches[0]] + batches:  # the graphed run spends its 2 warm-up steps on batches[0]
        loss = sh.loss_fn(b, m_e, sde, likelihood_weighting=False, rng="philox")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        eager.append(loss.item())

    torch.manual_seed(0)
    m_g = Unet(**kw).cuda().train()
    sh.manual_seed(99)
    step = GraphedTrainStep(m_g, sde, batches[0], lr=5e-4, warmup=2)
    got = [step(b).item() for b in batches]
    # eager[0:2] used batches[0] (= the graphed warm-up steps); compare from the third step on
    for a, b in zip(got, eager[2:]):
        assert abs(a - b) <= 2e-3 * abs(b), (got, eager)
    # Adam's normalised update turns split-K summation-order noise on near-zero gradients into O(lr) differences on
    # single elements, so compare the tensors in norm
    num = sum(((pe.double() - pg.double()) ** 2).sum() for pe, pg in zip(m_e.parameters(), m_g.parameters()))
    den = sum((pe.double() ** 2).sum() for pe in m_e.parameters())
    assert (num / den).sqrt().item() < 5e-3
    assert step.launches_per_step > 100
    # graph replays advance only the device counter; state_dict() still reports the true number of steps
    assert all(int(st["step"]) == 2 + len(batches) for st in step.opt.state_dict()["state"].values())


def test_update_ema_matches_reference_semantics():
    """utils.py:79-90: decay=0 copies the model (…_cond.py:674), then ema = ema*decay + p*(1-decay)."""
    from score_based_multimodal_autoencoder_b200.optim import update_ema
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(70, 300), torch.nn.Linear(300, 5)).cuda()
    e = torch.nn.Sequential(torch.nn.Linear(70, 300), torch.nn.Linear(300, 5)).cuda()
    update_ema(e, a, decay=0)
    for pe, pa in zip(e.parameters(), a.parameters()):
        assert torch.equal(pe, pa)
    ref = [p.detach().clone() for p in e.parameters()]
    for it in range(3):
        with torch.no_grad():
            for p in a.parameters():
                p.add_(torch.randn_like(p) * 0.1)
            for r, p in zip(ref, a.parameters()):
                r.mul_(0.99).add_(p, alpha=0.01)
        v0 = next(e.parameters())._version
        update_ema(e, a, decay=0.99)
        assert next(e.parameters())._version > v0
    for pe, r in zip(e.parameters(), ref):
        assert torch.allclose(pe, r, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("bucket_mb", [0.25, 1024.0])
def test_data_parallel_wrapper_on_one_gpu_leaves_the_plain_gradients(bucket_mb):
    """DataParallelScoreNet with world size 1 runs the whole gradient-sink path of the hand-written backward (slots of
    the flat bucket buffer, deferred weight-gradient unpacks, bucket launches) without a collective: its gradients must
    equal the plain model's.  (Round 2: gradients that were not produced in their slot were copied into the bucket
    buffer BEFORE their deferred unpack had run -- only the 2-GPU check saw it, and that test is skipped on a 1-GPU box.)"""
    from score_based_multimodal_autoencoder_b200 import distributed as D
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    torch.manual_seed(0)
    model = Unet(dim=32, channels=5, dim_mults=(1, 2, 2, 2)).cuda().train()
    sde = sh.VPSDE(1.0, 5.0, 8)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(16, 5, 8, 8, generator=g).cuda()
    u = torch.rand(16, generator=g).cuda()
    zz = torch.randn(16, 5, 8, 8, generator=g).cuda()
    sh.loss_fn(z, model, sde, likelihood_weighting=False, u=u, z=zz).backward()
    ref = [p.grad.detach().clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    ddp = D.DataParallelScoreNet(model, bucket_mb=bucket_mb)
    try:
        sh.loss_fn(z, ddp, sde, likelihood_weighting=False, u=u, z=zz).backward()
        torch.cuda.synchronize()
        assert len(ddp.reducer.launched) == len(ddp.reducer.buckets)
        # same kernels: only the order of the fp32 atomics (split-K, GroupNorm parameter sums) differs between passes
        num = sum(((p.grad.double() - r.double()) ** 2).sum() for p, r in zip(model.parameters(), ref))
        den = sum((r.double() ** 2).sum() for r in ref)
        assert (num / den).sqrt().item() < 1e-3
        worst = max((((p.grad.double() - r.double()).norm() / (r.double().norm() + 1e-30)).item(), n)
                    for (n, p), r in zip(model.named_parameters(), ref))
        assert worst[0] < 1e-2, worst   # an unwritten gradient is off by O(1)
    finally:
        del model._grad_sink
