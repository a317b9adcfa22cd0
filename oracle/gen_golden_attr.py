"""Generate tests/golden/attr_ae.pt: latents and reconstructions of the UNMODIFIED reference `CelebAAttrNewBN` /
`CelebAAttrNewBNAE` (h_vae_model.py:712-899, the attribute modality of train_lat_celebhq_unet_cont2.py:459-461) in eval
mode with deterministic weights and BatchNorm statistics, and check oracle/vae_oracle.py against them.

Run in the build container only:  python -m oracle.gen_golden_attr"""
from __future__ import annotations

import os
import sys
import types

import torch

from . import vae_oracle as vo
from .det_weights import fill_autoencoder_state_dict
from .gen_golden import OUT, REF


def main():
    for mod in ("torchvision", "torchvision.models"):
        if mod not in sys.modules:
            try:
                __import__(mod)
            except Exception:  # noqa: BLE001
                sys.modules[mod] = types.ModuleType(mod)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import h_vae_model as hv
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(9, 18, generator=g) > 0.5).float()      # binary attribute vectors
    zz = torch.randn(9, 256, generator=g)
    out = {"x": x, "zz": zz, "size_z": 256}
    for name, ref in (("vae", hv.CelebAAttrNewBN(256)), ("ae", hv.CelebAAttrNewBNAE(256))):
        sd0 = ref.state_dict()
        shapes = {k: tuple(v.shape) for k, v in sd0.items() if v.dtype.is_floating_point}
        sd = fill_autoencoder_state_dict(shapes, gain=1.6)
        full = dict(sd0)
        full.update(sd)
        ref.load_state_dict(full)
        ref.eval()
        with torch.no_grad():
            z = ref.encoder(x)
            z, logvar = (z, None) if name == "ae" else z
            rec = ref.decoder(zz)
        mu_o, lv_o = vo.attr_encode(sd, x)
        rec_o = vo.attr_decode(sd, zz)
        e = [((mu_o - z).abs().max() / z.abs().max()).item(), ((rec_o - rec).abs().max() / rec.abs().max()).item()]
        var_z = ((z - z.mean(0)).norm() / z.norm()).item()
        var_r = ((rec - rec.mean(0)).norm() / rec.norm()).item()
        assert var_z > 0.2 and var_r > 0.2, (var_z, var_r)     # the nets must not be dead: outputs depend on the inputs
        if logvar is not None:
            e.append(((lv_o - logvar).abs().max() / logvar.abs().max()).item())
        print(f"{name}: latent {tuple(z.shape)}, reconstruction {tuple(rec.shape)}, oracle rel-max {max(e):.2e}")
        assert max(e) < 1e-5
        out[name] = {"shapes": shapes, "z": z.clone(), "rec": rec.clone(), "logvar": None if logvar is None else logvar.clone()}
    path = os.path.join(OUT, "attr_ae.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
