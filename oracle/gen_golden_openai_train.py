"""Generate tests/golden/unet_openai_train.pt: DSM loss + gradients of the UNMODIFIED reference `UNetModel`
(unet_openai.py:361-575, z-conditioned, dropout 0) under the reference's own `loss_fn` (sde_helper2.py:152-186).
The reference's shipped `loss_fn` has no `z_cond` argument (train_lat_celebhq_unet_cont2_cond.py passes one anyway), so
the conditioning code is bound into the score_fn closure, which is what that call means.

A second fixture, unet_openai_train_dropout.pt, holds the same for the net the reference actually trains
(train_lat_celebhq_unet_cont2_cond.py:651-653: dropout = 0.1, train() mode): `nn.Dropout.forward` of the unmodified
reference is fed the masks the B200 path draws (oracle/philox.py: Philox4x32-10 keyed by seed, draw = index of the
ResBlock in execution order), since the reference's own torch-RNG masks cannot be reproduced by another generator.

Run in the build container only:  python -m oracle.gen_golden_openai_train
"""
from __future__ import annotations

import os

import torch

from .det_weights import fill_state_dict
from .gen_golden import OUT, NoiseFeed, import_reference

KW = dict(in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1, attention_resolutions=(2,), dropout=0.0,
          channel_mult=(1, 2, 2), num_heads=2, use_z=True, z_dim=16)


DROP_SEED = 0x1234ABCD5678


def dropout_case(sh, uoa):
    import numpy as np
    from torch import nn

    from .philox import dropout_keep
    kw = dict(KW, dropout=0.1)
    torch.manual_seed(0)
    ref = uoa.UNetModel(**kw)
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ref.load_state_dict(fill_state_dict(shapes))
    ref.train()
    g = torch.Generator().manual_seed(78)
    B = 6
    batch = torch.randn(B, 3, 8, 8, generator=g)
    u = torch.rand(B, generator=g)
    zn = torch.randn(B, 3, 8, 8, generator=g)
    zc = torch.randn(B, 16, generator=g)
    calls = [0]
    kept = []

    def fed_dropout(self, x):
        assert self.training and self.p == 0.1
        b, c, h, w = x.shape
        keep = dropout_keep(DROP_SEED, calls[0], b * h * w, c, self.p).reshape(b, h, w, c)
        calls[0] += 1
        kept.append(float(keep.mean()))
        return x * torch.from_numpy(np.ascontiguousarray(keep.transpose(0, 3, 1, 2))).to(x.dtype) / (1.0 - self.p)

    orig = nn.Dropout.forward
    nn.Dropout.forward = fed_dropout
    try:
        sde = sh.VPSDE(0.1, 20.0, 1000)
        with NoiseFeed([zn], [u]).patched():
            loss = sh.loss_fn(batch, lambda x, t: ref(x, t, z=zc), sde, reduce_mean=True, likelihood_weighting=False)
        loss.backward()
    finally:
        nn.Dropout.forward = orig
    grads = {k: {"norm": p.grad.norm().clone(), "head": p.grad.flatten()[:256].clone()}
             for k, p in ref.named_parameters() if p.grad is not None}
    gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in ref.parameters() if p.grad is not None))
    # the same forward WITHOUT dropout, to show the masks matter (the GPU test must not pass by ignoring them)
    ref.eval()
    with NoiseFeed([zn], [u]).patched():
        loss_nodrop = sh.loss_fn(batch, lambda x, t: ref(x, t, z=zc), sde, reduce_mean=True, likelihood_weighting=False)
    print(f"dropout 0.1: {calls[0]} masks (kept {min(kept):.3f}..{max(kept):.3f}); loss {loss.item():.6f} "
          f"(eval-mode loss {loss_nodrop.item():.6f}), grad norm {gnorm.item():.6f}")
    path = os.path.join(OUT, "unet_openai_train_dropout.pt")
    torch.save({"kwargs": kw, "shapes": shapes, "batch": batch, "u": u, "z": zn, "zc": zc, "seed": DROP_SEED,
                "n_masks": calls[0], "loss": loss.detach().clone(), "loss_eval_mode": loss_nodrop.detach().clone(),
                "grads": grads, "grad_norm": gnorm.detach().clone()}, path)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    sh, _, uoa = import_reference()
    dropout_case(sh, uoa)
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
