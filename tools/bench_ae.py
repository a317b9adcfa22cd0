"""Measure the residual autoencoders either side of the score-model path on one GPU: PolyMNIST
configuration of train_poly_unet_cont.py:548-560 (32x32x3 images, size_z 64), encode and decode at batch B, and the
parity numbers against the reference golden.  Prints one JSON line.  Usage: python tools/bench_ae.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.det_weights import fill_autoencoder_state_dict  # noqa: E402  (parity numbers only: the checker, not the thing measured)
from score_based_multimodal_autoencoder_b200 import _lib as L  # noqa: E402
from score_based_multimodal_autoencoder_b200.h_vae_model_copy import ResAE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.load(os.path.join(ROOT, "tests", "golden", "res_ae.pt"))
m = ResAE(g["enc"], g["dec"], g["size_in"], g["size_z"], g["img_ch"])
sd = fill_autoencoder_state_dict(g["ae"]["shapes"], gain=1.0)
full = dict(m.state_dict())
full.update(sd)
m.load_state_dict(full)
m = m.cuda().eval()


def rel(a, b):
    return ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (L.launch_count() - n0) // iters


z_par = rel(m.encoder(g["x"].cuda()), g["ae"]["z"])
r_par = rel(m.decoder(g["ae"]["z"].cuda()), g["ae"]["rec"])
x = torch.rand(B, g["img_ch"], g["size_in"], g["size_in"], device="cuda")
z = torch.randn(B, g["size_z"], device="cuda")
enc_ms, enc_l = timed(lambda: m.encoder(x))
dec_ms, dec_l = timed(lambda: m.decoder(z))
print(json.dumps({"model": "ResAE PolyMNIST (enc [(64,64,64,2),(64,128,128,2),(128,256,256,2)], 32x32x3, size_z 64), eval",
                  "batch": B, "encode_ms": round(enc_ms, 3), "encode_images_per_sec": round(B / enc_ms * 1e3),
                  "encode_launches": enc_l, "decode_ms": round(dec_ms, 3),
                  "decode_images_per_sec": round(B / dec_ms * 1e3), "decode_launches": dec_l,
                  "parity_vs_reference_golden": {"latent_rel_l2": float(f"{z_par:.3e}"),
                                                 "reconstruction_rel_l2": float(f"{r_par:.3e}")}}))
