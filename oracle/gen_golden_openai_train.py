"""Generate tests/golden/unet_openai_train.pt: DSM loss + gradients of the UNMODIFIED reference `UNetModel`
(unet_openai.py:361-575, z-conditioned, dropout 0) under the reference's own `loss_fn` (sde_helper2.py:152-186).
The reference's shipped `loss_fn` has no `z_cond` argument (train_lat_celebhq_unet_cont2_cond.py passes one anyway), so
the conditioning code is bound into the score_fn closure, which is what that call means.

Run in the build container only:  python -m oracle.gen_golden_openai_train
"""
from __future__ import annotations

import os

import torch

from .det_weights import fill_state_dict
from .gen_golden import OUT, NoiseFeed, import_reference

KW = dict(in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1, attention_resolutions=(2,), dropout=0.0,
          channel_mult=(1, 2, 2), num_heads=2, use_z=True, z_dim=16)


def main():
    sh, _, uoa = import_reference()
    torch.manual_seed(0)
    ref = uoa.UNetModel(**KW)
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ref.load_state_dict(fill_state_dict(shapes))
    ref.train()
    g = torch.Generator().manual_seed(77)
    B = 6
    batch = torch.randn(B, 3, 8, 8, generator=g)
    u = torch.rand(B, generator=g)
    zn = torch.randn(B, 3, 8, 8, generator=g)
    zc = torch.randn(B, 16, generator=g)
    cases = []
    for with_z in (True, False):
        sde = sh.VPSDE(0.1, 20.0, 1000)
        ref.zero_grad()
        score_fn = (lambda x, t: ref(x, t, z=zc)) if with_z else (lambda x, t: ref(x, t))
        with NoiseFeed([zn], [u]).patched():
            loss = sh.loss_fn(batch, score_fn, sde, reduce_mean=True, likelihood_weighting=False)
        loss.backward()
        grads = {k: {"norm": p.grad.norm().clone(), "head": p.grad.flatten()[:256].clone()}
                 for k, p in ref.named_parameters() if p.grad is not None}
        gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in ref.parameters() if p.grad is not None))
        cases.append({"with_z": with_z, "loss": loss.detach().clone(), "grads": grads, "grad_norm": gnorm.detach().clone(),
                      "no_grad": [k for k, p in ref.named_parameters() if p.grad is None]})
        print(f"with_z={with_z}: loss {loss.item():.6f}, grad norm {gnorm.item():.6f}, {len(grads)} parameter gradients")
    path = os.path.join(OUT, "unet_openai_train.pt")
    torch.save({"kwargs": KW, "shapes": shapes, "batch": batch, "u": u, "z": zn, "zc": zc, "cases": cases}, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
