"""HBM roofline of the fused sampler-step kernels at 64k Poly latents (bench.py's `roofline_sampler_kernels` leg alone),
for A/B runs of SBM_SAMPLER_RESERVE / SBM_NOISE_BLOCKS_PER_SM.  Usage: python tools/time_sampler_kernels.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
r = bench.sampler_kernel_roofline(sh, sh.VPSDE(1.0, 5.0, 100), torch.device("cuda"), batch=batch)
pk, _ = bench.peaks()
knobs = " ".join(k + "=" + v for k, v in os.environ.items() if k.startswith("SBM_")) or "defaults"
print(json.dumps({"knobs": knobs, "us_per_pc_step": r["us_per_pc_step"], "GBps": r["achieved"],
                  "frac": r["achieved"] / pk["hbm_gbs"]}))
