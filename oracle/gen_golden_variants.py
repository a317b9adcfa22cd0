"""Generate tests/golden/unet_variants.pt: forward outputs, DSM loss and gradients of the UNMODIFIED reference score
nets on the constructor paths NO shipped command uses (they are inside the reference's module API all the same):

  * `Unet(use_convnext=False)`               -- ResnetBlock / Block, unet_model.py:49-90
  * `UNetModel(num_classes=K)` + `y`         -- label embedding added to the time embedding, unet_openai.py:417-419, 561-564
  * `UNetModel(use_scale_shift_norm=True)`   -- unet_openai.py:257-260, 296-300

(`UNetModel(conv_resample=False)` is not here: the reference itself cannot construct it -- unet_openai.py:209 calls
`avg_pool_nd(stride)`, which builds nn.AvgPool2d() without a kernel size and raises TypeError; main() checks that.)

Each case: eval-mode forward on seeded inputs, then `loss_fn` (sde_helper2.py:152-186) + backward with the noise fed in.
Weights come from oracle/det_weights.py (per-name seeds), so the fixture stores shapes, not tensors.

Run in the build container only:  python -m oracle.gen_golden_variants
"""
from __future__ import annotations

import os

import torch

from .det_weights import fill_state_dict
from .gen_golden import OUT, NoiseFeed, import_reference

UNET_KW = dict(dim=32, channels=3, dim_mults=(1, 2, 2), use_convnext=False, resnet_block_groups=8)
OPENAI_BASE = dict(in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1, attention_resolutions=(2,),
                   dropout=0.0, channel_mult=(1, 2, 2), num_heads=2)
OPENAI_CASES = {
    "openai_num_classes": dict(OPENAI_BASE, num_classes=5),
    "openai_scale_shift_norm": dict(OPENAI_BASE, use_scale_shift_norm=True),
    "openai_classes_scale_shift_z": dict(OPENAI_BASE, num_classes=5, use_scale_shift_norm=True, use_z=True, z_dim=16),
}


def _grads(ref):
    grads = {k: {"norm": p.grad.norm().clone(), "head": p.grad.flatten()[:256].clone()}
             for k, p in ref.named_parameters() if p.grad is not None}
    gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in ref.parameters() if p.grad is not None))
    return grads, gnorm.detach().clone(), [k for k, p in ref.named_parameters() if p.grad is None]


def _case(sh, ref, kw, *, size, seed, call):
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ref.load_state_dict(fill_state_dict(shapes))
    g = torch.Generator().manual_seed(seed)
    B = 6
    x = torch.randn(B, 3, size, size, generator=g)
    t = torch.rand(B, generator=g) * 0.98 + 0.01
    u = torch.rand(B, generator=g)
    zn = torch.randn(B, 3, size, size, generator=g)
    zc = torch.randn(B, 16, generator=g)
    y = torch.randint(0, 5, (B,), generator=g)
    ref.eval()
    with torch.no_grad():
        out = call(ref, x, t, zc, y)
    ref.train()
    ref.zero_grad()
    sde = sh.VPSDE(0.1, 20.0, 1000)
    with NoiseFeed([zn], [u]).patched():
        loss = sh.loss_fn(x, lambda a, b: call(ref, a, b, zc, y), sde, reduce_mean=True, likelihood_weighting=False)
    loss.backward()
    grads, gnorm, no_grad = _grads(ref)
    print(f"  forward |out| {out.norm().item():.4f}; loss {loss.item():.6f}; grad norm {gnorm.item():.6f}; "
          f"{len(grads)} parameter gradients")
    return {"kwargs": kw, "shapes": shapes, "x": x, "t": t, "u": u, "z": zn, "zc": zc, "y": y, "out": out.clone(),
            "loss": loss.detach().clone(), "grads": grads, "grad_norm": gnorm, "no_grad": no_grad}


def main():
    sh, um, uoa = import_reference()
    fix = {}
    try:
        uoa.UNetModel(**dict(OPENAI_BASE, conv_resample=False))
        raise SystemExit("the reference constructed UNetModel(conv_resample=False): add a case for it")
    except TypeError as e:
        print("reference UNetModel(conv_resample=False):", type(e).__name__, e)
    print("unet_resnet_blocks", UNET_KW)
    torch.manual_seed(0)
    fix["unet_resnet_blocks"] = _case(sh, um.Unet(**UNET_KW), UNET_KW, size=8, seed=91,
                                      call=lambda m, x, t, zc, y: m(x, t))
    for i, (name, kw) in enumerate(OPENAI_CASES.items()):
        print(name, kw)
        torch.manual_seed(0)
        with_y, with_z = kw.get("num_classes") is not None, kw.get("use_z", False)

        def call(m, x, t, zc, y, with_y=with_y, with_z=with_z):
            return m(x, t, **(dict(z=zc) if with_z else {}), **(dict(y=y) if with_y else {}))

        fix[name] = _case(sh, uoa.UNetModel(**kw), kw, size=8, seed=92 + i, call=call)
    path = os.path.join(OUT, "unet_variants.pt")
    torch.save(fix, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
