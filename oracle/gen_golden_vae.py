"""Generate tests/golden/res_ae.pt: latents and reconstructions of the UNMODIFIED reference `ResAE` / `ResVAE`
(h_vae_model_copy.py) in eval mode with deterministic non-trivial weights and BatchNorm running statistics, for the
PolyMNIST configuration of train_poly_unet_cont.py:548-560 (32x32 inputs, three down-sampling RBlocks) at reduced batch.
`torchvision` (imported but unused by that file) is absent here: an empty module stands in for it.

Run in the build container only:  python -m oracle.gen_golden_vae"""
from __future__ import annotations

import os
import sys
import types

import torch

from . import vae_oracle as vo
from .det_weights import fill_autoencoder_state_dict, structured_images
from .gen_golden import OUT, REF

ENC = [(64, 64, 64, 2), (64, 128, 128, 2), (128, 256, 256, 2)]
DEC = [(256, 128, 128, 2), (128, 128, 64, 2), (64, 64, 64, 2)]
SIZE_IN, SIZE_Z, IMG_CH = 32, 64, 3


def main():
    if "torchvision" not in sys.modules:
        sys.modules["torchvision"] = types.ModuleType("torchvision")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import h_vae_model_copy as hv
    out = {"enc": ENC, "dec": DEC, "size_in": SIZE_IN, "size_z": SIZE_Z, "img_ch": IMG_CH}
    g = torch.Generator().manual_seed(11)
    x = structured_images(5, IMG_CH, SIZE_IN, 11)
    zz = torch.randn(5, SIZE_Z, generator=g)            # decoder inputs that differ a lot from each other
    for name, cls in (("ae", hv.ResAE), ("vae", hv.ResVAE)):
        torch.manual_seed(0)
        ref = cls(ENC, DEC, SIZE_IN, SIZE_Z, IMG_CH)
        sd0 = ref.state_dict()
        shapes = {k: tuple(v.shape) for k, v in sd0.items() if v.dtype.is_floating_point}
        sd = fill_autoencoder_state_dict(shapes, gain=1.0)
        full = dict(sd0)
        full.update(sd)
        ref.load_state_dict(full)
        ref.eval()
        with torch.no_grad():
            if name == "ae":
                z = ref.encoder(x)
                logvar = None
            else:
                z, logvar = ref.encoder(x)
            rec = ref.decoder(z)
            rec_zz = ref.decoder(zz)
            z_o = vo.ae_encode(sd, x, ENC)
            rec_o = vo.ae_decode(sd, z, ENC, DEC, SIZE_IN)
            rec_zz_o = vo.ae_decode(sd, zz, ENC, DEC, SIZE_IN)
        e1 = ((z_o - z).abs().max() / z.abs().max()).item()
        e2 = max(((rec_o - rec).abs().max() / rec.abs().max()).item(),
                 ((rec_zz_o - rec_zz).abs().max() / rec_zz.abs().max()).item())
        var_z = ((z - z.mean(0)).norm() / z.norm()).item()
        var_r = ((rec_zz - rec_zz.mean(0)).norm() / rec_zz.norm()).item()
        print(f"{name}: latent {tuple(z.shape)} oracle rel-max {e1:.2e}; reconstruction {tuple(rec.shape)} rel-max {e2:.2e}; "
              f"input-dependent part of the latents {var_z:.2f}, of the reconstructions {var_r:.2f}")
        assert e1 < 1e-5 and e2 < 1e-5 and var_z > 0.03 and var_r > 0.1
        if logvar is not None:
            lv_o = vo.res_encoder(sd, x, ENC)[1]
            assert ((lv_o - logvar).abs().max() / logvar.abs().max()).item() < 1e-5
        out[name] = {"shapes": shapes, "z": z.clone(), "rec": rec.clone(), "rec_zz": rec_zz.clone(),
                     "logvar": None if logvar is None else logvar.clone()}
    out["x"] = x
    out["zz"] = zz
    # ---- family "N" (CelebA-HQ image modality, train_lat_celebhq_unet_cont2.py:427-431) at reduced size
    enc_n, dec_n, size_n, z_n = [(64, 128, 128, 4), (128, 256, 256, 4)], [(256, 256, 128, 4), (128, 128, 64, 4)], 64, 256
    xn = structured_images(3, IMG_CH, size_n, 12)
    zzn = torch.randn(3, z_n, generator=g)
    out["N"] = {"enc": enc_n, "dec": dec_n, "size_in": size_n, "size_z": z_n, "x": xn, "zz": zzn}
    for name, cls in (("aen", hv.ResAEN), ("vaen", hv.ResVAEN)):
        torch.manual_seed(0)
        ref = cls(enc_n, dec_n, size_n, z_n, IMG_CH)
        sd0 = ref.state_dict()
        shapes = {k: tuple(v.shape) for k, v in sd0.items() if v.dtype.is_floating_point}
        sd = fill_autoencoder_state_dict(shapes, gain=1.0)
        full = dict(sd0)
        full.update(sd)
        ref.load_state_dict(full)
        ref.eval()
        with torch.no_grad():
            z = ref.encoder(xn)
            z = z if name == "aen" else z[0]
            rec = ref.decoder(z)
            rec_zz = ref.decoder(zzn)
            z_o = vo.ae_encode(sd, xn, enc_n, family="N")
            rec_o = vo.ae_decode(sd, z, enc_n, dec_n, size_n, family="N")
            rec_zz_o = vo.ae_decode(sd, zzn, enc_n, dec_n, size_n, family="N")
        e1 = ((z_o - z).abs().max() / z.abs().max()).item()
        e2 = max(((rec_o - rec).abs().max() / rec.abs().max()).item(),
                 ((rec_zz_o - rec_zz).abs().max() / rec_zz.abs().max()).item())
        var_z = ((z - z.mean(0)).norm() / z.norm()).item()
        var_r = ((rec_zz - rec_zz.mean(0)).norm() / rec_zz.norm()).item()
        print(f"{name}: latent {tuple(z.shape)} oracle rel-max {e1:.2e}; reconstruction {tuple(rec.shape)} rel-max {e2:.2e}; "
              f"input-dependent part of the latents {var_z:.2f}, of the reconstructions {var_r:.2f}")
        assert e1 < 1e-5 and e2 < 1e-5 and var_z > 0.03 and var_r > 0.02
        out["N"][name] = {"shapes": shapes, "z": z.clone(), "rec": rec.clone(), "rec_zz": rec_zz.clone()}
    path = os.path.join(OUT, "res_ae.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
