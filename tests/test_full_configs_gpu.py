"""Parity at the reference's REAL configurations and step counts:

  * N-step drift of the conditional PC sampler with the bf16 score net and injected noise, against the fp32 CPU oracle:
    PolyMNIST net, VPSDE(1, 5), N = 100 (train_poly.sh:17, train_poly_unet_cont.py:843) and the CelebA SDE
    VPSDE(0.1, 20), N = 1000 (train_cel.sh:11) with a narrow net; the curve is printed at intermediate step counts.
    STATED BOUND: rel-L2 of the sampler state <= 2e-2 after N = 100 steps and after N = 1000 steps (measured on B200:
    5.0e-3 and 6.6e-3; SURVEY.md App. E: the reference's own autocast(bf16) run drifts 4.2e-3 after 100 steps);
  * `UNetModel` at the full z-conditioned CelebA configuration (train_lat_celebhq_unet_cont2_cond.py:648-653,
    model_channels 128, channel_mult (1,2,4,8), z_dim 512): rows of a 256-sample batch against the oracle;
  * `loss_fn(likelihood_weighting=True, im_sample=True)` (sde_helper2.py:129-150, 164-165, 177-179) against a golden of
    the unmodified reference;
  * the CelebA-HQ image autoencoder (3 stages, 512 channels, 128x128) and the 1-channel mask autoencoder
    (train_lat_celebhq_unet_cont2.py:427-452) against the fp32 oracle (itself pinned by tests/golden/res_ae.pt).
"""
import pytest
import torch

from oracle import sde_oracle as so
from oracle import unet_oracle as uo
from oracle import vae_oracle as vo
from oracle.det_weights import fill_autoencoder_state_dict, fill_state_dict, structured_images
from tests.util import golden, rel_l2

pytestmark = pytest.mark.gpu


def _sh():
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    return sh


def _drift_curve(kw, M, D, a, b, N, B, given, mods, marks, seed):
    """GPU sampler state after k steps (k in marks) vs the oracle's trace; same weights, same injected noise."""
    sh = _sh()
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    m = Unet(**kw)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd)
    m = m.cuda().eval()
    sde = sh.VPSDE(a, b, N)
    spec = so.SdeSpec("vp", a, b, N)
    g = torch.Generator().manual_seed(seed)
    z0 = torch.randn(B, M, D, D, generator=g)
    x0 = torch.randn(B, M, D, D, generator=g)
    npred = torch.randn(N, B, M, D, D, generator=g)
    ncorr = torch.randn(N, 1, B, M, D, D, generator=g)
    mask = [k in given for k in mods]
    score_fn = lambda x, t: uo.unet_forward(sd, x, t, dim=kw["dim"], dim_mults=kw["dim_mults"])
    with torch.no_grad():
        ref_out, trace = so.pc_sampler(spec, score_fn, x0, npred, ncorr, z_obs=z0, obs_mask=mask, return_trace=True)
    curve = {}
    miss = [i for i, on in enumerate(mask) if not on]
    for k in marks:
        if k == N:  # the sampler's output: x_mean of the last step on the missing channels, clean observed channels
            out = sh.cond_sampler(z0.cuda(), given, mods, m, sde, x_init=x0.cuda(), noise_pred=npred.cuda(),
                                  noise_corr=ncorr.cuda())
            curve[k] = rel_l2(out, ref_out)
        else:       # the sampler STATE after k steps against the oracle's trace
            _, st = sh.pc_sampler(x0.cuda(), m, sde, z_obs=z0.cuda(), obs_mask=sh._obs_mask_from(given, mods),
                                  noise_pred=npred.cuda(), noise_corr=ncorr.cuda(), num_steps=k, return_state=True)
            curve[k] = rel_l2(st[:, miss], trace[k - 1][:, miss])
    return curve


def test_drift_after_100_steps_poly_net():
    """BASELINE configs[0]'s net and SDE at the reference's real step count."""
    curve = _drift_curve(dict(dim=64, channels=5, dim_mults=(1, 2, 2, 2)), 5, 8, 1.0, 5.0, 100, 8, "0", "01234",
                         (1, 10, 25, 50, 100), seed=31)
    print("Poly-64 net, VPSDE(1,5), N=100, B=8: drift (rel-L2 vs fp32 oracle) " +
          ", ".join(f"{k}: {v:.3e}" for k, v in curve.items()))
    assert all(v < 2e-2 for v in curve.values()), curve


def test_drift_after_1000_steps_celeba_sde():
    """The CelebA SDE (beta 0.1..20, N = 1000, 3 modalities of 16x16) with a narrow three-level net (dim 32) so that the
    fp32 CPU oracle finishes its 2000 forwards in about half a minute."""
    curve = _drift_curve(dict(dim=32, channels=3, dim_mults=(1, 2, 2)), 3, 16, 0.1, 20.0, 1000, 2, "0", "012",
                         (1, 10, 100, 500, 1000), seed=32)
    print("dim-32 three-level net on the CelebA latent, VPSDE(0.1,20), N=1000, B=2: drift (rel-L2 vs fp32 oracle) " +
          ", ".join(f"{k}: {v:.3e}" for k, v in curve.items()))
    assert all(v < 2e-2 for v in curve.values()), curve


def test_unetmodel_full_celeba_config_rows_vs_oracle():
    """226 M-parameter z-conditioned score net of train_lat_celebhq_unet_cont2_cond.py:648-653 at batch 256."""
    from score_based_multimodal_autoencoder_b200.unet_openai import UNetModel
    kw = dict(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(),
              dropout=0.1, channel_mult=(1, 2, 4, 8), num_heads=1, use_z=True, z_dim=512)
    m = UNetModel(**kw)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd)
    m = m.cuda().eval()
    n_params = sum(p.numel() for p in m.parameters())
    g = torch.Generator().manual_seed(41)
    B = 256
    x = torch.randn(B, 3, 16, 16, generator=g)
    t = torch.rand(B, generator=g) * 0.999 + 1e-3
    z = torch.randn(B, 512, generator=g)
    pick = torch.tensor([0, 131, 255])
    with torch.no_grad():
        y = m(x.cuda(), t.cuda(), z=z.cuda())
        ref = uo.unet_openai_forward(sd, x[pick], t[pick], z=z[pick], model_channels=128, num_res_blocks=2,
                                     attention_resolutions=(), channel_mult=(1, 2, 4, 8), num_heads=1)
    err = rel_l2(y[pick.cuda()], ref)
    print(f"UNetModel(128, (1,2,4,8), z_dim 512): {n_params / 1e6:.1f} M parameters, batch {B}, rows {pick.tolist()}: "
          f"rel-L2 vs oracle = {err:.3e}")
    assert y.shape == (B, 3, 16, 16) and torch.isfinite(y).all()
    assert err < 1.5e-2


def test_loss_fn_importance_sampled_time_branch_matches_reference_golden(monkeypatch):
    sh = _sh()
    g = golden("dsm_loss_is.pt")
    batch, z, w0 = g["batch"].cuda(), g["z"].cuda(), g["w"]
    for c in g["cases"]:
        sde = sh.VPSDE(c["a"], c["b"], c["N"])
        # the reference draws the quantile with torch.distributions.Uniform (CPU torch.rand underneath): feed the same
        feed = [g["u01"].clone()]
        real_rand = torch.rand
        monkeypatch.setattr(torch, "rand", lambda *a, **k: feed.pop(0) if feed else real_rand(*a, **k))
        w = w0.clone().cuda().requires_grad_(True)
        score = lambda xx, tt: torch.einsum("oc,bchw->bohw", w, xx) * (1.0 + tt[:, None, None, None])
        loss = sh.loss_fn(batch, score, sde, reduce_mean=c["reduce_mean"], likelihood_weighting=True, im_sample=True,
                          z=z)
        monkeypatch.setattr(torch, "rand", real_rand)
        loss.backward()
        assert not feed
        assert abs(loss.item() - c["loss"].item()) <= 2e-5 * abs(c["loss"].item()), (loss.item(), c["loss"].item())
        assert rel_l2(w.grad, c["grad_w"]) < 2e-5


@pytest.mark.parametrize("name,enc,dec,img_ch,B", [
    ("image", [(64, 128, 128, 2), (128, 256, 256, 2), (256, 512, 512, 2)],
     [(512, 512, 256, 2), (256, 256, 128, 2), (128, 128, 64, 2)], 3, 2),
    ("mask", [(64, 128, 128, 4), (128, 256, 256, 4)], [(256, 256, 128, 4), (128, 128, 64, 4)], 1, 3),
])
def test_celeba_autoencoders_full_config_vs_oracle(name, enc, dec, img_ch, B):
    """ResAEN at the CelebAMask-HQ sizes (128x128 inputs, size_z 256; train_lat_celebhq_unet_cont2.py:427-452)."""
    from score_based_multimodal_autoencoder_b200 import h_vae_model_copy as hv
    size_in, size_z = 128, 256
    m = hv.ResAEN(enc, dec, size_in, size_z, img_ch)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    sd = fill_autoencoder_state_dict(shapes, gain=1.0)
    full = dict(m.state_dict())
    full.update(sd)
    m.load_state_dict(full)
    m = m.cuda().eval()
    x = structured_images(B, img_ch, size_in, 21)
    zz = torch.randn(B, size_z, generator=torch.Generator().manual_seed(22))
    with torch.no_grad():
        z_ref = vo.ae_encode(sd, x, enc, family="N")
        rec_ref = vo.ae_decode(sd, zz, enc, dec, size_in, family="N")
    z = m.encoder(x.cuda())
    rec = m.decoder(zz.cuda())
    e_z, e_r = rel_l2(z, z_ref), rel_l2(rec, rec_ref)
    dep = lambda a, b: ((a.double().cpu() - b.double()).norm() / (b.double() - b.double().mean(0, keepdim=True)).norm()).item()
    print(f"CelebA {name} ResAEN 128x128: latent rel-L2 {e_z:.3e} (input-dependent part {dep(z, z_ref):.3e}), "
          f"reconstruction {e_r:.3e} (input-dependent part {dep(rec, rec_ref):.3e})")
    assert z.shape == (B, size_z) and rec.shape == (B, img_ch, size_in, size_in)
    assert e_z < 2e-2 and e_r < 2e-2
    assert dep(z, z_ref) < 8e-2 and dep(rec, rec_ref) < 8e-2
