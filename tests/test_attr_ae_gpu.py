"""GPU parity of the attribute-modality autoencoders (score_based_multimodal_autoencoder_b200/h_vae_model.py, SURVEY.md
8f-1) against the golden outputs of the unmodified reference `CelebAAttrNewBN` / `CelebAAttrNewBNAE` in eval mode
(tests/golden/attr_ae.pt) and against the fp32 oracle at another batch.  bf16 GEMM operands: rel-L2 <= 2e-2."""
import pytest
import torch

from oracle import vae_oracle as vo
from oracle.det_weights import fill_autoencoder_state_dict
from tests.util import golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-2        # whole output
TOL_VAR = 6e-2    # input-dependent part only (the golden's nets are alive: > 20 % of the output depends on the input)


def rel_var(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b - b.mean(0, keepdim=True)).norm()).item()


def _build(name, g):
    from score_based_multimodal_autoencoder_b200 import h_vae_model as hm
    m = hm.CelebAAttrNewBN(g["size_z"]) if name == "vae" else hm.CelebAAttrNewBNAE(g["size_z"])
    sd = fill_autoencoder_state_dict(g[name]["shapes"], gain=1.6)
    full = dict(m.state_dict())
    full.update(sd)
    m.load_state_dict(full)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("name", ["vae", "ae"])
def test_attr_autoencoder_matches_reference_golden(name):
    g = golden("attr_ae.pt")
    m, sd = _build(name, g)
    x = g["x"].cuda()
    if name == "vae":
        z, logvar = m.encoder(x)
        assert rel_l2(logvar, g[name]["logvar"]) < TOL
    else:
        z = m.encoder(x)
    rec = m.decoder(g["zz"].cuda())
    e_z, e_r = rel_l2(z, g[name]["z"]), rel_l2(rec, g[name]["rec"])
    v_z, v_r = rel_var(z, g[name]["z"]), rel_var(rec, g[name]["rec"])
    print(f"attr {name}: latent rel-L2 {e_z:.3e} (input-dependent part {v_z:.3e}), reconstruction {e_r:.3e} ({v_r:.3e})")
    assert z.shape == g[name]["z"].shape and rec.shape == g[name]["rec"].shape and rec.dtype == torch.float32
    assert e_z < TOL and e_r < TOL and v_z < TOL_VAR and v_r < TOL_VAR
    # another batch size against the oracle; per-sample independence; train() mode is rejected
    gen = torch.Generator().manual_seed(5)
    x2 = (torch.rand(300, 18, generator=gen) > 0.5).float()
    z2 = m.encoder(x2.cuda())
    z2 = z2[0] if name == "vae" else z2
    assert rel_l2(z2, vo.attr_encode(sd, x2)[0]) < TOL
    z1 = m.encoder(x2[:2].cuda())
    z1 = z1[0] if name == "vae" else z1
    assert rel_l2(z1, z2[:2]) < 1e-5
    m.train()
    with pytest.raises(NotImplementedError):
        m.decoder(z2)
