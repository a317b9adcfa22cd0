"""GPU parity of the drop-in residual autoencoders (score_based_multimodal_autoencoder_b200/h_vae_model_copy.py,
SURVEY.md 8f-1) against the golden outputs of the unmodified reference `ResAE` / `ResVAE` in eval mode
(tests/golden/res_ae.pt) and against the fp32 CPU oracle at another batch size.  bf16 GEMM operands and bf16
activations between blocks, fp32 accumulation: rel-L2 <= 2e-2 of the fp32 reference."""
import pytest
import torch

from oracle import vae_oracle as vo
from oracle.det_weights import fill_autoencoder_state_dict, structured_images
from tests.util import golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-2        # whole output
TOL_VAR = 8e-2    # input-dependent part only (output minus its batch mean): a broken data path shows up here as ~1


def rel_var(a, b):
    """error relative to the INPUT-DEPENDENT part of the reference (b minus its mean over the batch)."""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b - b.mean(0, keepdim=True)).norm()).item()


def _build(name, g):
    from score_based_multimodal_autoencoder_b200 import h_vae_model_copy as hv
    cls = {"ae": hv.ResAE, "vae": hv.ResVAE, "aen": hv.ResAEN, "vaen": hv.ResVAEN}[name]
    m = cls(g["enc"], g["dec"], g["size_in"], g["size_z"], g.get("img_ch", 3))
    sd = fill_autoencoder_state_dict(g[name]["shapes"], gain=1.0)
    full = dict(m.state_dict())
    full.update(sd)
    m.load_state_dict(full)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("name", ["ae", "vae"])
def test_res_autoencoder_matches_reference_golden(name):
    g = golden("res_ae.pt")
    m, _ = _build(name, g)
    x = g["x"].cuda()
    if name == "ae":
        z = m.encoder(x)
    else:
        z, logvar = m.encoder(x)
        assert rel_l2(logvar, g[name]["logvar"]) < TOL
    rec = m.decoder(g[name]["z"].cuda())
    rec_zz = m.decoder(g["zz"].cuda())
    e_z, e_r, e_rz = rel_l2(z, g[name]["z"]), rel_l2(rec, g[name]["rec"]), rel_l2(rec_zz, g[name]["rec_zz"])
    v_z, v_rz = rel_var(z, g[name]["z"]), rel_var(rec_zz, g[name]["rec_zz"])
    print(f"{name}: latent rel-L2 {e_z:.3e} (input-dependent part {v_z:.3e}), reconstruction {e_r:.3e}, "
          f"reconstruction of random latents {e_rz:.3e} (input-dependent part {v_rz:.3e})")
    assert z.shape == g[name]["z"].shape and rec.shape == g[name]["rec"].shape and rec.dtype == torch.float32
    assert e_z < TOL and e_r < TOL and e_rz < TOL and v_z < TOL_VAR and v_rz < TOL_VAR
    # whole round trip through the module's own forward
    out = m(x)
    out = out if name == "ae" else out[0]
    assert out.shape == x.shape and torch.isfinite(out).all()


@pytest.mark.parametrize("name", ["aen", "vaen"])
def test_res_autoencoder_celeba_variants_match_reference_golden(name):
    """ResAEN / ResVAEN (GELU blocks, bilinear up-sampling by 4, LeakyReLU(0.1) stem, sigmoid output) against the
    unmodified reference (tests/golden/res_ae.pt, key "N")."""
    n = golden("res_ae.pt")["N"]
    m, _ = _build(name, n)
    z = m.encoder(n["x"].cuda())
    z = z if name == "aen" else z[0]
    rec = m.decoder(n[name]["z"].cuda())
    rec_zz = m.decoder(n["zz"].cuda())
    e_z, e_r, e_rz = rel_l2(z, n[name]["z"]), rel_l2(rec, n[name]["rec"]), rel_l2(rec_zz, n[name]["rec_zz"])
    v_z, v_rz = rel_var(z, n[name]["z"]), rel_var(rec_zz, n[name]["rec_zz"])
    print(f"{name}: latent rel-L2 {e_z:.3e} (input-dependent part {v_z:.3e}), reconstruction {e_r:.3e}, "
          f"reconstruction of random latents {e_rz:.3e} (input-dependent part {v_rz:.3e})")
    assert z.shape == n[name]["z"].shape and rec.shape == n[name]["rec"].shape
    assert e_z < TOL and e_r < TOL and e_rz < TOL and v_z < TOL_VAR and v_rz < TOL_VAR
    assert rec.min() >= 0 and rec.max() <= 1


@pytest.mark.parametrize("case", [(2, 4, 4, 64, 4, torch.float32), (3, 8, 8, 24, 2, torch.bfloat16),
                                  (2, 2, 2, 16, 4, torch.float32), (2, 1, 1, 8, 2, torch.float32)])
def test_gelu_bilinear_resample_kernel(case):
    """sbm_act_resample: exact GELU then bilinear up-sampling (nn.Upsample(mode='bilinear')) / average pooling."""
    import torch.nn.functional as F
    from score_based_multimodal_autoencoder_b200 import ops
    from score_based_multimodal_autoencoder_b200 import h_vae_model_copy as hv
    B, H, W, Cc, rate, dt = case
    gen = torch.Generator().manual_seed(B * 10 + H + Cc)
    x = torch.randn(B, Cc, H, W, generator=gen).cuda()
    xn = torch.full((B, H, W, ops.pad8(Cc)), float("nan"), device="cuda", dtype=dt)
    xn[..., :Cc] = x.permute(0, 2, 3, 1).to(dt)
    xr = xn[..., :Cc].float().permute(0, 3, 1, 2)
    act = F.gelu(xr.double())
    up = F.interpolate(act, scale_factor=rate, mode="bilinear").permute(0, 2, 3, 1)
    out = hv.lrelu_resample(xn, Cc, 0.0, hv.MODE_BILINEAR, rate, act=hv.ACT_GELU)
    torch.cuda.synchronize()
    assert out.shape[1:3] == up.shape[1:3]
    assert (out[..., :Cc].double() - up).abs().max().item() <= 5e-3 * up.abs().max().item() + 1e-6
    assert (out[..., Cc:] == 0).all()
    if H % 2 == 0:
        pool = F.avg_pool2d(act, 2).permute(0, 2, 3, 1)
        o2 = hv.lrelu_resample(xn, Cc, 0.0, hv.MODE_AVGPOOL, 2, act=hv.ACT_GELU)
        assert (o2[..., :Cc].double() - pool).abs().max().item() <= 5e-3 * pool.abs().max().item() + 1e-6


def test_res_autoencoder_matches_oracle_at_another_batch_and_rejects_train_mode():
    g = golden("res_ae.pt")
    m, sd = _build("ae", g)
    x = structured_images(37, g["img_ch"], g["size_in"], 3)
    zz = torch.randn(37, g["size_z"], generator=torch.Generator().manual_seed(4))
    z_ref = vo.ae_encode(sd, x, g["enc"])
    z = m.encoder(x.cuda())
    rec = m.decoder(zz.cuda())
    rec_ref = vo.ae_decode(sd, zz, g["enc"], g["dec"], g["size_in"])
    assert rel_l2(z, z_ref) < TOL and rel_l2(rec, rec_ref) < TOL
    assert rel_var(z, z_ref) < TOL_VAR and rel_var(rec, rec_ref) < TOL_VAR
    # per-sample independence (eval-mode BatchNorm is a fixed affine map)
    assert rel_l2(m.encoder(x[:3].cuda()), z[:3]) < 1e-5
    m.train()
    with pytest.raises(NotImplementedError):
        m.encoder(x.cuda())
    m.eval()
    from score_based_multimodal_autoencoder_b200 import _lib as L
    with pytest.raises(L.SbmError):
        m.encoder(x)


@pytest.mark.parametrize("case", [(3, 8, 8, 40, 0, 1, torch.float32), (2, 16, 16, 64, 1, 2, torch.float32),
                                  (2, 8, 8, 24, 1, 4, torch.bfloat16), (3, 4, 4, 64, 2, 2, torch.float32),
                                  (2, 2, 2, 16, 2, 4, torch.bfloat16)])
def test_lrelu_resample_kernel(case):
    """sbm_lrelu_resample against torch: LeakyReLU then AvgPool2d / nearest up-sampling, both input types, NCHW output."""
    import torch.nn.functional as F
    from score_based_multimodal_autoencoder_b200 import ops
    from score_based_multimodal_autoencoder_b200.h_vae_model_copy import lrelu_resample
    B, H, W, Cc, mode, rate, dt = case
    gen = torch.Generator().manual_seed(B * 100 + H + Cc)
    x = torch.randn(B, Cc, H, W, generator=gen).cuda()
    ld = ops.pad8(Cc)
    xn = torch.full((B, H, W, ld), float("nan"), device="cuda", dtype=dt)
    xn[..., :Cc] = x.permute(0, 2, 3, 1).to(dt)
    xr = xn[..., :Cc].float().permute(0, 3, 1, 2)
    ref = F.leaky_relu(xr, 0.2)
    if mode == 1:
        ref = F.avg_pool2d(ref, rate)
    elif mode == 2:
        ref = F.interpolate(ref, scale_factor=rate, mode="nearest")
    out = lrelu_resample(xn, Cc, 0.2, mode, rate)
    torch.cuda.synchronize()
    assert out.shape[1:3] == ref.shape[2:] and out.dtype == torch.bfloat16
    assert torch.equal(out[..., :Cc].float(), ref.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    assert (out[..., Cc:] == 0).all()
    if mode == 0:
        o2 = lrelu_resample(xn, Cc, 0.2, nchw=True)
        assert torch.equal(o2, ref)


def test_images_to_latents_to_conditional_samples_to_images():
    """The reference's conditional generation loop end to end on the CUDA path (train_poly_unet_cont.py:421-471, 257-268):
    encode every modality with its frozen autoencoder, stack the latents [B, M, 8, 8], run the conditional PC sampler
    with modality 0 observed, decode the sampled latents of the missing modalities."""
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    g = golden("res_ae.pt")
    B, mods = 12, "01234"
    gen = torch.Generator().manual_seed(9)
    aes = {}
    for i, mkey in enumerate(mods):
        torch.manual_seed(100 + i)
        aes[mkey], _ = _build("ae", g)
    imgs = {mkey: torch.rand(B, g["img_ch"], g["size_in"], g["size_in"], generator=gen).cuda() for mkey in mods}
    z = {mkey: aes[mkey].encoder(imgs[mkey]) for mkey in mods}                       # {mod: [B, 64]}
    assert all(v.shape == (B, g["size_z"]) for v in z.values())
    torch.manual_seed(0)
    score = Unet(dim=32, channels=5, dim_mults=(1, 2, 2)).cuda().eval()
    sde = sh.VPSDE(1.0, 5.0, 20)
    sh.manual_seed(4)
    out = sh.cond_sampler(z, "0", mods, score, sde, num_steps=5)
    assert out.shape == (B, len(mods), 8, 8)                                         # stacked latent, channel = modality
    assert torch.equal(out[:, 0].reshape(B, -1), z["0"])                             # the observed modality is returned clean
    for i, mkey in enumerate(mods):
        if i == 0:
            continue
        zi = out[:, i].reshape(B, g["size_z"])
        assert torch.isfinite(zi).all()
        rec = aes[mkey].decoder(zi)
        assert rec.shape == imgs[mkey].shape and rec.dtype == torch.float32 and torch.isfinite(rec).all()
