"""Measure the z-conditioned `UNetModel` path of train_lat_celebhq_unet_cont2_cond.py:648-674 on one GPU:
conditional PC sampling (eval mode, `z_cond`), the DSM training step with dropout 0.1 (eager and as one CUDA graph) and
the EMA update.  Prints one JSON line.  Usage: python tools/bench_openai.py [sample_batch] [train_batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402
from score_based_multimodal_autoencoder_b200.optim import FusedAdam, GraphedTrainStep, update_ema  # noqa: E402
from score_based_multimodal_autoencoder_b200.unet_openai import UNetModel  # noqa: E402

BS = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
BT = int(sys.argv[2]) if len(sys.argv) > 2 else 256
KW = dict(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(), dropout=0.1,
          channel_mult=(1, 2, 4, 8), num_heads=1, use_z=True, z_dim=512)
GF_FWD = 5.4677  # GFLOP per sample per forward
dev = torch.device("cuda")
sde = sh.VPSDE(0.1, 20.0, 1000)


def fresh():
    torch.manual_seed(0)
    m = UNetModel(**KW).to(dev)
    with torch.no_grad():  # the reference zero-initialises some convs: re-randomise so every kernel does real work
        for p in m.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.02)
    return m


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"model": "UNetModel(3, 128, 3, 2, (), dropout=0.1, channel_mult=(1,2,4,8), use_z=True, z_dim=512)", "latent": [3, 16, 16]}
g = torch.Generator().manual_seed(1)
# ---- conditional PC sampling, modality 0 observed, z-conditioned
m = fresh().eval()
z_obs = torch.randn(BS, 3, 16, 16, generator=g).to(dev)
zc = torch.randn(BS, 512, generator=g).to(dev)
with torch.no_grad():
    sh.cond_sampler(z_obs, "0", "012", m, sde, num_steps=2, z_cond=zc)
    K = 10
    ms = timed(lambda: sh.cond_sampler(z_obs, "0", "012", m, sde, num_steps=K, z_cond=zc), 1) / K
out["sampling"] = {"batch": BS, "ms_per_pc_step_eager": round(ms, 3), "samples_per_sec_N1000": round(BS / ms, 2)}
with torch.no_grad():
    t = torch.full((BS,), 0.5, device=dev)
    x = torch.randn(BS, 3, 16, 16, device=dev)
    for _ in range(2):
        m(x, t, z=zc)
    fwd = timed(lambda: m(x, t, z=zc), 10)
out["sampling"].update({"ms_per_forward_eager": round(fwd, 3), "model_tflops": round(GF_FWD * BS / fwd, 1)})   # GFLOP / ms = TFLOP/s
del m
# ---- DSM training with dropout 0.1 (train mode), batch BT
batch = torch.randn(BT, 3, 16, 16, generator=g).to(dev)
zct = torch.randn(BT, 512, generator=g).to(dev)
m = fresh().train()
ema = fresh()
opt = FusedAdam(m.parameters(), lr=5e-5)


def eager_step():
    loss = sh.loss_fn(batch, m, sde, likelihood_weighting=False, rng="philox", z_cond=zct)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    eager_step()
eager = timed(eager_step, 5)
update_ema(ema, m, decay=0.999)
ema_ms = timed(lambda: update_ema(ema, m, decay=0.999), 10)
del opt
m2 = fresh().train()
step = GraphedTrainStep(m2, sde, batch, lr=5e-5, warmup=3, loss_kwargs=dict(likelihood_weighting=False, z_cond=zct))
for _ in range(2):
    step(batch)
graphed = timed(lambda: step(batch), 10)
loss = step(batch).item()
out["dsm_train"] = {"batch": BT, "dropout": 0.1, "ms_per_step_eager": round(eager, 3), "ms_per_step_graph": round(graphed, 3),
                    "steps_per_sec": round(1e3 / graphed, 2), "launches_per_step": step.launches_per_step,
                    "model_tflops": round(3 * GF_FWD * BT / graphed, 1), "ema_update_ms": round(ema_ms, 3),
                    "loss": round(loss, 4)}
step.close()
print(json.dumps(out))
