"""CPU tests (no GPU): the oracle restatement against the golden vectors produced by the real reference
(oracle/gen_golden.py), the host-side SDE objects, state_dict compatibility and the C-ABI symbol table."""
import ctypes
import os
import re

import pytest
import torch

from oracle import sde_oracle as so
from oracle import unet_oracle as uo
from oracle.det_weights import fill_autoencoder_state_dict, fill_state_dict
from tests.util import ROOT, golden, rel_l2


@pytest.mark.parametrize("name", ["unet_poly", "unet_cel"])
def test_oracle_unet_matches_reference_golden(name):
    g = golden(name + ".pt")
    sd = fill_state_dict(g["shapes"])
    with torch.no_grad():
        y = uo.unet_forward(sd, g["x"], g["t"], dim=g["kwargs"]["dim"], dim_mults=g["kwargs"]["dim_mults"])
    assert rel_l2(y, g["y"]) < 1e-5


def test_oracle_openai_unet_matches_reference_golden():
    g = golden("unet_openai.pt")
    sd = fill_state_dict(g["shapes"])
    kw = dict(model_channels=32, num_res_blocks=1, attention_resolutions=(2,), channel_mult=(1, 2, 2), num_heads=2)
    with torch.no_grad():
        assert rel_l2(uo.unet_openai_forward(sd, g["x"], g["t"], z=g["z"], **kw), g["y"]) < 1e-5
        assert rel_l2(uo.unet_openai_forward(sd, g["x"], g["t"], **kw), g["y_noz"]) < 1e-5


def test_oracle_sde_objects_and_host_classes():
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    cls = {"vp": sh.VPSDE, "subvp": sh.subVPSDE, "ve": sh.VESDE}
    for c in golden("sde_objects.pt"):
        spec = so.SdeSpec(c["kind"], c["a"], c["b"], c["N"])
        d, gdiff = so.sde_coeffs(spec, c["x"], c["t"])
        m, s = so.marginal_prob(spec, c["x"], c["t"])
        assert torch.equal(d, c["drift"]) and torch.allclose(gdiff, c["diffusion"], rtol=1e-6)
        assert torch.equal(m, c["mean"]) and torch.equal(s, c["std"])
        assert torch.allclose(so.prior_logp(spec, c["x"]), c["prior_logp"], rtol=1e-6)
        # the drop-in SDE classes (host-side tensor expressions) reproduce the reference bit for bit on CPU
        sde = cls[c["kind"]](c["a"], c["b"], c["N"])
        d2, g2 = sde.sde(c["x"], c["t"])
        m2, s2 = sde.marginal_prob(c["x"], c["t"])
        assert torch.equal(d2, c["drift"]) and torch.allclose(g2, c["diffusion"], rtol=1e-6)
        assert torch.equal(m2, c["mean"]) and torch.equal(s2, c["std"])
        assert torch.allclose(sde.prior_logp(c["x"]), c["prior_logp"], rtol=1e-6)
        if "disc_f" in c:
            f, G = sde.discretize(c["x"], c["t"])
            assert torch.allclose(f, c["disc_f"], rtol=1e-6, atol=1e-7) and torch.allclose(G, c["disc_G"], rtol=1e-6)
            fo, Go = so.discretize(spec, c["x"], c["t"])
            assert torch.equal(fo, c["disc_f"]) and torch.equal(Go, c["disc_G"])
        if "alphas" in c:
            assert torch.equal(sde.alphas, c["alphas"])
            assert torch.equal(sde.sqrt_1m_alphas_cumprod, c["sqrt_1m_alphas_cumprod"])
            assert torch.equal(spec.alphas(), c["alphas"])
        assert isinstance(sde, sh.SDE) and sde.T == 1 and sde.N == c["N"]


def test_oracle_reverse_diffusion_predictor_matches_reference_golden():
    """oracle discretize / rd_predictor_step == the unmodified reference's sde.reverse(...).discretize based update
    (oracle/gen_golden_rd.py), all three SDE kinds, SDE and ODE; the host classes' discretize() agree too."""
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    cls = {"vp": sh.VPSDE, "subvp": sh.subVPSDE, "ve": sh.VESDE}
    for c in golden("rd_predictor.pt"):
        spec = so.SdeSpec(c["kind"], c["a"], c["b"], c["N"])
        f, G = so.discretize(spec, c["x"], c["t"])
        assert torch.equal(f, c["disc_f"]) and torch.equal(G, c["disc_G"])
        f2, G2 = cls[c["kind"]](c["a"], c["b"], c["N"]).discretize(c["x"], c["t"])
        assert torch.allclose(f2, c["disc_f"], rtol=1e-6, atol=1e-7) and torch.allclose(G2, c["disc_G"], rtol=1e-6)
        for pf, key in ((False, "sde"), (True, "ode")):
            xn, xm = so.rd_predictor_step(spec, c["x"], c["t"], c["score"], c["z"], pf)
            assert torch.equal(xn, c[key]["x"]) and torch.equal(xm, c[key]["x_mean"])


def test_oracle_sampler_steps_and_loops():
    g = golden("sampler_steps.pt")
    kind, a, b, N = g["sde"]
    spec = so.SdeSpec(kind, a, b, N)
    x, t = g["x"], g["t"]
    xp, xm = so.em_predictor_step(spec, x, t, g["score"], g["z_pred"])
    assert rel_l2(xp, g["pred_x"]) < 1e-6 and rel_l2(xm, g["pred_mean"]) < 1e-6
    xc, xcm = so.corrector_step(spec, x, t, g["score"], g["z_corr"], g["target_snr"])
    assert rel_l2(xc, g["corr_x"]) < 1e-6 and rel_l2(xcm, g["corr_mean"]) < 1e-6
    net = golden("unet_poly.pt")
    sd = fill_state_dict(net["shapes"])
    score_fn = lambda xx, tt: uo.unet_forward(sd, xx, tt, dim=net["kwargs"]["dim"], dim_mults=net["kwargs"]["dim_mults"])
    with torch.no_grad():
        for loop in g["loops"]:
            mask = [m in loop["given"] for m in g["mods"]]
            out = so.pc_sampler(spec, score_fn, g["loop_z0"], g["loop_npred"], g["loop_ncorr"], z_obs=g["loop_z0"],
                                obs_mask=mask, noise_obs=loop["noise_obs"], predictor_first=loop["predictor_first"],
                                num_steps=g["loop_steps"])
            assert rel_l2(out, loop["out"]) < 1e-5, loop["given"]


def test_oracle_dsm_loss():
    g = golden("dsm_loss.pt")
    net = golden("unet_poly.pt")
    sd = fill_state_dict(net["shapes"])
    score_fn = lambda xx, tt: uo.unet_forward(sd, xx, tt, dim=net["kwargs"]["dim"], dim_mults=net["kwargs"]["dim_mults"])
    with torch.no_grad():
        for c in g["cases"]:
            spec = so.SdeSpec(c["kind"], c["a"], c["b"], c["N"])
            loss = so.dsm_loss(spec, g["batch"], score_fn, g["u"], g["z"], reduce_mean=c["reduce_mean"],
                               likelihood_weighting=c["likelihood_weighting"])
            assert abs(loss.item() - c["loss"].item()) <= 1e-5 * abs(c["loss"].item())


def test_state_dict_schema_matches_reference():
    """Checkpoints of the reference must load: identical key order, names and shapes."""
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    for name in ("unet_poly", "unet_cel"):
        g = golden(name + ".pt")
        m = Unet(**g["kwargs"])
        sd = m.state_dict()
        assert list(sd.keys()) == list(g["shapes"].keys())
        assert {k: tuple(v.shape) for k, v in sd.items()} == g["shapes"]
        m.load_state_dict(fill_state_dict(g["shapes"]))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sbmae_b200.h")).read()
    names = set(re.findall(r"\b(sbm_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    lib_path = os.path.join(ROOT, "score_based_multimodal_autoencoder_b200", "csrc", "libsbmae_b200.so")
    if not os.path.exists(lib_path):
        from score_based_multimodal_autoencoder_b200.build import build
        build()
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    lib.sbm_last_error.restype = ctypes.c_char_p
    assert lib.sbm_version() >= 100


def test_product_path_has_no_cpu_fallback():
    from score_based_multimodal_autoencoder_b200 import _lib as L
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    m = Unet(dim=32, channels=5, dim_mults=(1, 2))
    with pytest.raises(L.SbmError):
        m(torch.zeros(1, 5, 8, 8), torch.ones(1))
    with pytest.raises(L.SbmError):
        sh.em_predictor(torch.zeros(1, 5, 8, 8), torch.ones(1), lambda x, t: x, sh.VPSDE())
    # the product package never imports the oracle
    import score_based_multimodal_autoencoder_b200 as pkg
    for fn in os.listdir(os.path.dirname(pkg.__file__)):
        if fn.endswith(".py"):
            src = open(os.path.join(os.path.dirname(pkg.__file__), fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_unet_openai_state_dict_schema_matches_reference_golden():
    """`UNetModel` keeps the reference's parameter names and shapes (checkpoints load unchanged)."""
    from score_based_multimodal_autoencoder_b200.unet_openai import UNetModel
    g = golden("unet_openai.pt")
    m = UNetModel(**g["kwargs"])
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in g["shapes"].items()}
    with pytest.raises(Exception):
        m(g["x"], g["t"])  # CPU tensors: no fallback


def test_oracle_philox_known_answers_and_dropout_mask():
    """oracle/philox.py against the Random123 known-answer vectors of Philox4x32-10 (kat_vectors: zero, all-ones and
    pi-digit counter / key), and the statistics of the dropout mask derived from it."""
    import numpy as np

    from oracle.philox import dropout_keep, philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(v) for v in got) == want
    keep = dropout_keep(0x1234ABCD5678, 3, 4096, 20, 0.1)
    assert keep.shape == (4096, 20) and abs(keep.mean() - 0.9) < 5e-3
    assert not np.array_equal(keep, dropout_keep(0x1234ABCD5678, 4, 4096, 20, 0.1))
    assert dropout_keep(1, 0, 64, 8, 0.0).all()


def test_docs_name_only_declared_entry_points_and_cover_all_of_them():
    """INTEGRATION.md is the binding guide: every `sbm_*` it (or DESIGN.md) names is declared in include/sbmae_b200.h,
    and every declared entry point appears in INTEGRATION.md's table."""
    hdr = open(os.path.join(ROOT, "include", "sbmae_b200.h")).read()
    known = set(re.findall(r"\b(sbm_[a-z0-9_]+)\b", hdr))
    declared = set(re.findall(r"\b(sbm_[a-z0-9_]+)\s*\(", hdr))
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for doc in ("INTEGRATION.md", "DESIGN.md"):
        named = set(re.findall(r"\b(sbm_[a-z0-9_]+)\b", open(os.path.join(ROOT, doc)).read()))
        assert not (named - known), (doc, sorted(named - known))
    assert not [n for n in declared if n not in integ], sorted(n for n in declared if n not in integ)


def test_oracle_residual_autoencoders_match_reference_golden():
    """oracle/vae_oracle.py (the encoders / decoders either side of the score-model path) against the
    unmodified reference ResAE / ResVAE in eval mode (tests/golden/res_ae.pt, oracle/gen_golden_vae.py)."""
    from oracle import vae_oracle as vo
    g = golden("res_ae.pt")
    for name in ("ae", "vae"):
        c = g[name]
        sd = fill_autoencoder_state_dict(c["shapes"], gain=1.0)
        mu, logvar = vo.res_encoder(sd, g["x"], g["enc"])
        assert torch.allclose(mu, c["z"], rtol=1e-5, atol=1e-6)
        if c["logvar"] is not None:
            assert torch.allclose(logvar, c["logvar"], rtol=1e-5, atol=1e-6)
        rec = vo.ae_decode(sd, c["z"], g["enc"], g["dec"], g["size_in"])
        assert rec.shape == g["x"].shape and torch.allclose(rec, c["rec"], rtol=1e-5, atol=1e-5)
        assert torch.allclose(vo.ae_decode(sd, g["zz"], g["enc"], g["dec"], g["size_in"]), c["rec_zz"], rtol=1e-5, atol=1e-5)
        # the fixtures are input-sensitive (a golden whose outputs barely depend on the inputs pins only the biases)
        assert ((c["z"] - c["z"].mean(0)).norm() / c["z"].norm()).item() > 0.03
        assert ((c["rec_zz"] - c["rec_zz"].mean(0)).norm() / c["rec_zz"].norm()).item() > 0.1
    # the CelebA-HQ variants ResAEN / ResVAEN (GELU blocks, bilinear up-sampling, sigmoid output): oracle pinned for the
    # next round's CUDA side
    n = g["N"]
    for name in ("aen", "vaen"):
        c = n[name]
        sd = fill_autoencoder_state_dict(c["shapes"], gain=1.0)
        assert torch.allclose(vo.ae_encode(sd, n["x"], n["enc"], family="N"), c["z"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(vo.ae_decode(sd, n["zz"], n["enc"], n["dec"], n["size_in"], family="N"), c["rec_zz"],
                              rtol=1e-5, atol=1e-6)
        rec = vo.ae_decode(sd, c["z"], n["enc"], n["dec"], n["size_in"], family="N")
        assert torch.allclose(rec, c["rec"], rtol=1e-5, atol=1e-6) and rec.min() >= 0 and rec.max() <= 1


def test_res_autoencoder_state_dict_schema_matches_reference_golden():
    """The drop-in ResAE / ResVAE expose the reference's parameter and BatchNorm-buffer names and shapes, so its
    checkpoints load (h_vae_model_copy.py:92-174)."""
    from score_based_multimodal_autoencoder_b200 import h_vae_model_copy as hv
    g = golden("res_ae.pt")
    for name, cls in (("ae", hv.ResAE), ("vae", hv.ResVAE)):
        m = cls(g["enc"], g["dec"], g["size_in"], g["size_z"], g["img_ch"])
        got = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
        assert got == g[name]["shapes"]
    n = g["N"]
    for name, cls in (("aen", hv.ResAEN), ("vaen", hv.ResVAEN)):
        m = cls(n["enc"], n["dec"], n["size_in"], n["size_z"], 3)
        got = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
        assert got == n[name]["shapes"]


def test_oracle_attribute_autoencoders_and_schema_match_reference_golden():
    """oracle/vae_oracle.py `attr_*` against the unmodified reference CelebAAttrNewBN / CelebAAttrNewBNAE
    (tests/golden/attr_ae.pt, oracle/gen_golden_attr.py), and the drop-in classes' state_dict schema."""
    from oracle import vae_oracle as vo
    from score_based_multimodal_autoencoder_b200 import h_vae_model as hm
    g = golden("attr_ae.pt")
    for name in ("vae", "ae"):
        c = g[name]
        sd = fill_autoencoder_state_dict(c["shapes"], gain=1.6)
        mu, logvar = vo.attr_encode(sd, g["x"])
        assert torch.allclose(mu, c["z"], rtol=1e-5, atol=1e-6)
        assert (logvar is None) == (c["logvar"] is None)
        if logvar is not None:
            assert torch.allclose(logvar, c["logvar"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(vo.attr_decode(sd, g["zz"]), c["rec"], rtol=1e-5, atol=1e-5)
        assert ((c["z"] - c["z"].mean(0)).norm() / c["z"].norm()).item() > 0.2      # not a dead ReLU net
        m = hm.CelebAAttrNewBN(g["size_z"]) if name == "vae" else hm.CelebAAttrNewBNAE(g["size_z"])
        got = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
        assert got == c["shapes"]


def test_only_tests_smoke_and_bench_import_the_oracle():
    """oracle/ is the checker: nothing under the package or tools/ may import it (bench.py: cpu_baseline / --impl
    reference only; __graft_entry__: build() compiles nothing from it, smoke() checks against it)."""
    for sub in ("score_based_multimodal_autoencoder_b200", "tools"):
        d = os.path.join(ROOT, sub)
        for fn in os.listdir(d):
            if fn.endswith(".py"):
                src = open(os.path.join(d, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, (sub, fn)


def test_autoencoder_host_plans_on_emulated_kernels_match_the_oracle():
    """The host side of the autoencoder plans (BatchNorm folding, weight permutations, slicing, call order) run against
    torch stand-ins of the kernels (tests/emulation.py) and compared with the fp32 oracle on input-sensitive weights:
    whole output and, separately, the input-dependent part of it (a broken data path shows up there as ~1)."""
    from oracle import vae_oracle as vo
    from oracle.det_weights import structured_images
    from score_based_multimodal_autoencoder_b200 import h_vae_model as hm
    from score_based_multimodal_autoencoder_b200 import h_vae_model_copy as hc
    from tests.emulation import emulated_kernels

    def rel_var(a, b):
        a, b = a.double(), b.double()
        return ((a - b).norm() / (b - b.mean(0, keepdim=True)).norm()).item()

    def load(m, gain):
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
        sd = fill_autoencoder_state_dict(shapes, gain)
        full = dict(m.state_dict())
        full.update(sd)
        m.load_state_dict(full)
        return m.eval(), sd

    poly = ([(64, 64, 64, 2), (64, 128, 128, 2), (128, 256, 256, 2)], [(256, 128, 128, 2), (128, 128, 64, 2), (64, 64, 64, 2)])
    with emulated_kernels(), torch.no_grad():
        for fam, cls, (enc, dec), size, zdim, nb in (("", hc.ResVAE, poly, 32, 64, 4),
                                                     ("N", hc.ResVAEN, ([(64, 128, 128, 4), (128, 256, 256, 4)],
                                                                        [(256, 256, 128, 4), (128, 128, 64, 4)]), 64, 256, 3)):
            m, sd = load(cls(enc, dec, size, zdim, 3), 1.0)
            x = structured_images(nb, 3, size, 11)
            zz = torch.randn(nb, zdim, generator=torch.Generator().manual_seed(3))
            mu, lv = m.encoder(x)
            rec = m.decoder(zz)
            mu_o, lv_o = vo.res_encoder(sd, x, enc, family=fam)
            rec_o = vo.ae_decode(sd, zz, enc, dec, size, family=fam)
            assert rel_l2(mu, mu_o) < 1e-2 and rel_l2(lv, lv_o) < 1e-2 and rel_l2(rec, rec_o) < 1e-2, fam
            assert rel_var(mu, mu_o) < 8e-2 and rel_var(rec, rec_o) < 8e-2, fam
        for m in (hm.CelebAAttrNewBN(256), hm.CelebAAttrNewBNAE(256)):
            m, sd = load(m, 1.6)
            x = (torch.rand(16, 18, generator=torch.Generator().manual_seed(5)) > 0.5).float()
            zz = torch.randn(16, 256, generator=torch.Generator().manual_seed(6))
            z = m.encoder(x)
            z = z[0] if isinstance(z, tuple) else z
            rec = m.decoder(zz)
            z_o, rec_o = vo.attr_encode(sd, x)[0], vo.attr_decode(sd, zz)
            assert rel_l2(z, z_o) < 1e-2 and rel_l2(rec, rec_o) < 1e-2
            assert rel_var(z, z_o) < 6e-2 and rel_var(rec, rec_o) < 6e-2
    # the real kernel wrappers are back: a CPU tensor is refused again
    from score_based_multimodal_autoencoder_b200 import _lib as L
    with pytest.raises(L.SbmError):
        hm.CelebAAttrNewBNAE(64).eval().decoder(torch.zeros(2, 64))


def test_checkpoint_container_round_trip_and_reference_compatibility(tmp_path):
    """train_lat_celebhq_unet_cont2.py:534-557, :480: {'epoch','model_state_dict','train_loss','val_loss','size_z'}."""
    import os
    from score_based_multimodal_autoencoder_b200 import eval_samplers as es
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    torch.manual_seed(0)
    m = Unet(dim=32, channels=3, dim_mults=(1, 2))
    p = str(tmp_path / "ck")
    es.save_checkpoint(p, m, epoch=3, train_loss=0.1, val_loss=0.2, size_z=256)
    ck = torch.load(p, map_location="cpu", weights_only=False)
    assert sorted(ck) == ["epoch", "model_state_dict", "size_z", "train_loss", "val_loss"]
    torch.manual_seed(1)
    m2 = Unet(dim=32, channels=3, dim_mults=(1, 2))
    assert es.load_checkpoint(p, m2) == {"epoch": 3, "train_loss": 0.1, "val_loss": 0.2, "size_z": 256}
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    if os.path.isdir("/root/reference"):      # build container: the file loads strictly into the reference's own module
        from oracle.gen_golden import import_reference
        _, um, _ = import_reference()
        ref = um.Unet(dim=32, channels=3, dim_mults=(1, 2))
        ref.load_state_dict(ck["model_state_dict"], strict=True)


def test_oracle_variants_match_reference_golden():
    """Oracle restatements of the constructor paths no shipped command uses (ResnetBlock Unet, class-conditional and
    scale-shift-norm UNetModel) against tests/golden/unet_variants.pt (the unmodified reference, run by
    oracle/gen_golden_variants.py)."""
    from oracle import unet_oracle as uo
    from oracle.det_weights import fill_state_dict
    fix = torch.load(os.path.join(ROOT, "tests", "golden", "unet_variants.pt"))
    for name, c in fix.items():
        kw, sd = c["kwargs"], fill_state_dict(c["shapes"])
        with torch.no_grad():
            if name == "unet_resnet_blocks":
                y = uo.unet_forward(sd, c["x"], c["t"], dim=kw["dim"], dim_mults=kw["dim_mults"], use_convnext=False,
                                    groups=kw["resnet_block_groups"])
            else:
                y = uo.unet_openai_forward(sd, c["x"], c["t"], model_channels=kw["model_channels"],
                                           num_res_blocks=kw["num_res_blocks"],
                                           attention_resolutions=kw["attention_resolutions"],
                                           channel_mult=kw["channel_mult"], num_heads=kw["num_heads"],
                                           z=c["zc"] if kw.get("use_z") else None,
                                           y=c["y"] if kw.get("num_classes") is not None else None)
        assert ((y - c["out"]).norm() / c["out"].norm()).item() < 1e-5, name


def test_variant_modules_mirror_the_reference_state_dict_and_errors():
    """Parameter names / shapes of the variant nets equal the reference's (fixture `shapes`), and
    UNetModel(conv_resample=False) raises the reference's own TypeError (unet_openai.py:209 builds nn.AvgPool2d()
    without a kernel size: the reference cannot construct that net either)."""
    import pytest
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    from score_based_multimodal_autoencoder_b200.unet_openai import UNetModel
    fix = torch.load(os.path.join(ROOT, "tests", "golden", "unet_variants.pt"))
    for name, c in fix.items():
        m = (Unet if name == "unet_resnet_blocks" else UNetModel)(**c["kwargs"])
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == c["shapes"], name
    with pytest.raises(TypeError, match="kernel_size"):
        UNetModel(in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1, attention_resolutions=(),
                  channel_mult=(1, 2), conv_resample=False)


def test_ctypes_struct_mirrors_match_the_header(tmp_path):
    """The `ctypes.Structure` mirrors in `_lib.py` are written by hand: compile a probe against include/sbmae_b200.h with
    gcc and compare sizeof and the offset of every field (an ABI drift here corrupts kernel arguments silently)."""
    import subprocess
    from score_based_multimodal_autoencoder_b200 import _lib as L
    pairs = {"sbm_conv_args": L.ConvArgs, "sbm_wgrad_args": L.WgradArgs, "sbm_pack_desc": L.PackDesc,
             "sbm_latent_shape": L.LatentShape, "sbm_sde": L.SdeC, "sbm_rng": L.Rng, "sbm_impute": L.Impute,
             "sbm_adam_tensor": L.AdamTensor, "sbm_ema_tensor": L.EmaTensor}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "sbmae_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), (cname, got[cname], ctypes.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_split_k_plan_host_logic():
    """Host-side plan of the split-K convolution (no kernel launch; without a device the library assumes the B200's 148
    SMs): which layers of the headline net are cut along K at which per-GPU batch, that the plan yields to the
    pixel-major tiling where that skips more work, and the workspace size it asks for."""
    from score_based_multimodal_autoencoder_b200 import _lib as L
    lib = L.lib()
    lib.sbm_conv_splitk(1)

    def plan(kind, k, b, h, cin, cout, nchw=0):
        a = L.ConvArgs()
        a.kind, a.kh, a.kw = kind, k, k
        a.batch, a.h, a.w = b, h, h
        a.cin, a.cout = cin, cout
        a.out_nchw = nchw
        a.ld_ws = (cout + 7) // 8 * 8
        return lib.sbm_conv_splitk_plan(ctypes.byref(a)), lib.sbm_conv_splitk_ws_elems(ctypes.byref(a))

    S1, S2 = L.CONV_S1, L.CONV_S2
    # 128 latents per GPU (the 8-GPU shard): 4x4 and 2x2 levels, both 3x3 convolutions of a ConvNeXt block
    assert plan(S1, 3, 128, 4, 512, 1024)[0] == 2      # 8 x 4 = 32 tiles on 74 SM pairs
    assert plan(S1, 3, 128, 4, 1024, 512)[0] == 4      # 16 tiles
    assert plan(S1, 3, 128, 2, 1024, 512)[0] == 8      # 4 tiles: capped at 8 slices
    # the stride-2 down-sampling convolutions (4x4 kernel): 8x8 -> 4x4 and 2x2 -> 1x1 (4 of 16 taps see real pixels)
    assert plan(S2, 4, 128, 8, 512, 512)[0] == 4
    assert plan(S2, 4, 256, 2, 512, 512)[0] == 4       # 32 K blocks / 8
    assert plan(S2, 4, 256, 2, 256, 512)[0] == 1       # 16 K blocks: too short to cut
    # 1024 latents: the 2x2 level would be 2 slices of 9 taps against 1 slice of 4 valid taps -> pixel-major wins
    assert plan(S1, 3, 1024, 2, 1024, 512)[0] == 1
    assert plan(S1, 3, 1024, 16, 256, 512)[0] == 1     # large layers: more tiles than SM pairs
    assert plan(S1, 1, 128, 4, 512, 1024)[0] == 1      # 1x1: K too short
    assert plan(S1, 3, 128, 4, 512, 128)[0] == 1       # narrow outputs run other tile shapes
    assert plan(S1, 3, 128, 4, 512, 1024, nchw=1)[0] == 1
    # workspace: slices x rows rounded up to whole 256-row tile pairs x row stride
    n, elems = plan(S1, 3, 100, 2, 512, 1024)          # M = 400 rows -> 512 per slab
    assert n == 8 and elems == 8 * 512 * 1024
    n, elems = plan(S2, 4, 128, 8, 512, 512)           # M = 2048
    assert elems == n * 2048 * 512
    lib.sbm_conv_splitk(0)
    try:
        assert plan(S1, 3, 128, 4, 512, 1024) == (1, 0)
    finally:
        lib.sbm_conv_splitk(1)
