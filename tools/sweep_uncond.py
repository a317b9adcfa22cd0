"""BASELINE configs[4]: large-batch UNCONDITIONAL predictor-corrector sampling sweep (uncond_sampler(..., pc=True)
semantics, sde_helper2.py:115-128: corrector then predictor per step), batch 4k-64k latents per GPU, PolyMNIST-5 and
CelebA score nets.  Times K replayed PC steps per point with CUDA events and prints one JSON line per point:
samples/s = batch / (N * s_per_step) for the N-step sample.  The score net walks batches beyond Unet.max_batch() in
slices (exact: the net has no cross-sample coupling); the fused sampler kernels take the whole batch.
Usage: python tools/sweep_uncond.py [poly|celeba|both] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402
from score_based_multimodal_autoencoder_b200.unet_model import Unet  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
GRAPH = os.environ.get("SBM_SWEEP_GRAPH", "1") == "1"   # replay the cached CUDA-graph step of the public entry point
NETS = {"poly": (dict(dim=64, channels=5, dim_mults=(1, 2, 2, 2)), (5, 8, 8), sh.VPSDE(1.0, 5.0, 1000)),
        "celeba": (dict(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)), (3, 16, 16), sh.VPSDE(0.1, 20.0, 1000))}
for name in (["poly", "celeba"] if which == "both" else [which]):
    kw, lat, sde = NETS[name]
    torch.manual_seed(0)
    m = Unet(**kw).cuda().eval()
    for B in (4096, 8192, 16384, 32768, 65536):
        sh.manual_seed(1234)
        x = sh.randn((B,) + lat, "cuda")
        with torch.no_grad():
            sh.pc_sampler(x, m, sde, pc=True, predictor_first=False, num_steps=2, use_graph=GRAPH)   # warm-up / capture
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = sh.pc_sampler(x, m, sde, pc=True, predictor_first=False, num_steps=K, use_graph=GRAPH)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(json.dumps({"workload": f"{name}_uncond_pc", "batch": B, "score_net_slice": min(B, m.max_batch(lat[1], lat[2])),
                          "ms_per_pc_step": round(ms, 3), "samples_per_sec_N1000": round(B / (sde.N * ms * 1e-3), 2),
                          "finite": bool(torch.isfinite(out).all()), "cuda_graph": GRAPH,
                          "knobs": " ".join(k + "=" + v for k, v in os.environ.items() if k.startswith("SBM_")) or "defaults",
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)
        del x, out
        sh.clear_graph_cache()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
