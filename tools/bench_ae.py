"""Measure the residual autoencoders either side of the score-model path on one GPU: PolyMNIST
configuration of train_poly_unet_cont.py:548-560 (32x32x3 images, size_z 64), encode and decode at batch B, and the
number of kernels per call (parity lives in tests/test_res_ae_gpu.py).  Prints one JSON line.
Usage: python tools/bench_ae.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from score_based_multimodal_autoencoder_b200 import _lib as L  # noqa: E402
from score_based_multimodal_autoencoder_b200.h_vae_model_copy import ResAE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ENC = [(64, 64, 64, 2), (64, 128, 128, 2), (128, 256, 256, 2)]      # train_poly_unet_cont.py:548-550
DEC = [(256, 128, 128, 2), (128, 128, 64, 2), (64, 64, 64, 2)]
g = {"img_ch": 3, "size_in": 32, "size_z": 64}
torch.manual_seed(0)
m = ResAE(ENC, DEC, g["size_in"], g["size_z"], g["img_ch"])
with torch.no_grad():                                               # non-trivial BatchNorm statistics
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
m = m.cuda().eval()


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


x = torch.rand(B, g["img_ch"], g["size_in"], g["size_in"], device="cuda")
z = torch.randn(B, g["size_z"], device="cuda")
enc_ms, enc_l = timed(lambda: m.encoder(x))
dec_ms, dec_l = timed(lambda: m.decoder(z))
print(json.dumps({"model": "ResAE PolyMNIST (enc [(64,64,64,2),(64,128,128,2),(128,256,256,2)], 32x32x3, size_z 64), eval",
                  "batch": B, "encode_ms": round(enc_ms, 3), "encode_images_per_sec": round(B / enc_ms * 1e3),
                  "encode_launches": enc_l, "decode_ms": round(dec_ms, 3),
                  "decode_images_per_sec": round(B / dec_ms * 1e3), "decode_launches": dec_l}))
