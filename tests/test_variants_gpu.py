"""GPU parity of the score-net constructor paths NO shipped command of the reference uses -- `Unet(use_convnext=False)`
(ResnetBlock, unet_model.py:49-90), `UNetModel(num_classes=K)` (unet_openai.py:417-419, 561-564) and
`UNetModel(use_scale_shift_norm=True)` (unet_openai.py:257-260, 296-300) -- against the real reference modules
(tests/golden/unet_variants.pt, made by oracle/gen_golden_variants.py): eval-mode forward, DSM loss, every parameter
gradient.  bf16 GEMM operands / fp32 accumulate: forward rel-L2 <= 1.5e-2, gradient bounds as for the shipped nets
(tests/test_unet_openai_gpu.py).  (CPU side: tests/test_oracle_cpu.py pins the oracle restatements to the same fixture.)"""
import pytest
import torch

from oracle.det_weights import fill_state_dict
from tests.util import golden, rel_l2

pytestmark = pytest.mark.gpu
BF16_NET_TOL = 1.5e-2
OPENAI_CASES = ["openai_num_classes", "openai_scale_shift_norm", "openai_classes_scale_shift_z"]


def _build(case, c):
    if case == "unet_resnet_blocks":
        from score_based_multimodal_autoencoder_b200.unet_model import Unet as Net
    else:
        from score_based_multimodal_autoencoder_b200.unet_openai import UNetModel as Net
    m = Net(**c["kwargs"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == c["shapes"]   # the reference's state_dict loads
    m.load_state_dict(fill_state_dict(c["shapes"]))
    return m.cuda()


def _call(case, c):
    kw = c["kwargs"]
    extra = {}
    if case != "unet_resnet_blocks":
        if kw.get("use_z"):
            extra["z"] = c["zc"].cuda()
        if kw.get("num_classes") is not None:
            extra["y"] = c["y"].cuda()
    return lambda m, x, t: m(x, t, **extra)


@pytest.mark.parametrize("case", ["unet_resnet_blocks"] + OPENAI_CASES)
def test_variant_forward_matches_reference_golden(case):
    c = golden("unet_variants.pt")[case]
    m = _build(case, c).eval()
    call = _call(case, c)
    with torch.no_grad():
        y = call(m, c["x"].cuda(), c["t"].cuda())
        y2 = call(m, c["x"].cuda(), c["t"].cuda())
    assert y.shape == c["out"].shape and y.dtype == torch.float32
    e = rel_l2(y, c["out"])
    print(f"{case}: forward rel-L2 vs reference = {e:.3e}")
    assert e < BF16_NET_TOL
    assert torch.equal(y, y2)   # no state leaks between calls (stat arenas, packed-weight caches)


@pytest.mark.parametrize("case", ["unet_resnet_blocks"] + OPENAI_CASES)
def test_variant_training_gradients_match_reference_golden(case):
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    c = golden("unet_variants.pt")[case]
    m = _build(case, c).train()
    call = _call(case, c)
    sde = sh.VPSDE(0.1, 20.0, 1000)
    loss = sh.loss_fn(c["x"].cuda(), lambda a, b: call(m, a, b), sde, reduce_mean=True, likelihood_weighting=False,
                      u=c["u"].cuda(), z=c["z"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - c["loss"].item()) <= 2e-2 * abs(c["loss"].item())
    params = dict(m.named_parameters())
    worst, worst_k, rels = 0.0, None, []
    for k, ref in c["grads"].items():
        got = params[k].grad
        assert got is not None, k
        if ref["norm"].item() < 1e-5 * c["grad_norm"].item():
            # a bias in front of a GroupNorm whose groups are single channels: exactly zero up to rounding noise
            assert got.norm().item() < 1e-4 * c["grad_norm"].item(), k
            continue
        head = got.flatten()[:256].float().cpu()
        rel = ((head - ref["head"]).norm() / (ref["head"].norm() + 1e-12)).item()
        nrel = abs(got.norm().item() - ref["norm"].item()) / (ref["norm"].item() + 1e-12)
        assert nrel <= 8e-2, (k, nrel)
        rels.append(rel)
        if rel > worst:
            worst, worst_k = rel, k
    for k in c["no_grad"]:
        assert params[k].grad is None or params[k].grad.abs().max().item() == 0.0, k
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None)).item()
    rels.sort()
    print(f"{case}: loss {loss.item():.5f} vs {c['loss'].item():.5f}; total grad norm {gn:.5e} vs "
          f"{c['grad_norm'].item():.5e}; head rel-L2 over {len(rels)} parameters: median {rels[len(rels) // 2]:.3e}, "
          f"p90 {rels[int(0.9 * len(rels))]:.3e}, worst {worst:.3e} ({worst_k})")
    assert abs(gn - c["grad_norm"].item()) <= 3e-2 * c["grad_norm"].item()
    assert rels[int(0.9 * len(rels))] <= 5e-2 and worst <= 0.2, (worst, worst_k)


def test_resnet_block_unet_in_the_sampler():
    """The ResnetBlock net behind the public sampler entry point: 2 conditional PC steps on the B200 path against the
    CPU oracle loop driven by the same net evaluated in fp32 (torch functional restatement of unet_model.py:49-90)."""
    from oracle import sde_oracle as so
    from oracle import unet_oracle as uo
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    c = golden("unet_variants.pt")["unet_resnet_blocks"]
    m = _build("unet_resnet_blocks", c).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    kw = c["kwargs"]
    score_fn = lambda x, t: uo.unet_forward(sd, x, t, dim=kw["dim"], dim_mults=kw["dim_mults"], use_convnext=False,
                                            groups=kw["resnet_block_groups"])
    with torch.no_grad():
        assert rel_l2(score_fn(c["x"], c["t"]), c["out"]) < 1e-5   # the oracle restatement is pinned by the golden
    g = torch.Generator().manual_seed(5)
    B, N, steps = 6, 10, 2
    z = torch.randn(B, 3, 8, 8, generator=g)
    npred = torch.randn(steps, B, 3, 8, 8, generator=g)
    ncorr = torch.randn(steps, 1, B, 3, 8, 8, generator=g)
    sde = sh.VPSDE(1.0, 5.0, N)   # N = 10 steps: the discrete betas must stay below 1 (alpha = 1 - beta > 0)
    out = sh.cond_sampler(z.cuda(), "0", "012", m, sde, x_init=z.cuda(), noise_pred=npred.cuda(),
                          noise_corr=ncorr.cuda(), num_steps=steps)
    with torch.no_grad():
        ref = so.pc_sampler(so.SdeSpec("vp", 1.0, 5.0, N), score_fn, z, npred, ncorr, z_obs=z,
                            obs_mask=[True, False, False], num_steps=steps)
    e = rel_l2(out, ref)
    print(f"ResnetBlock Unet, 2 conditional PC steps: rel-L2 vs oracle = {e:.3e}")
    assert e < 2e-2
