"""ORACLE helper (test infrastructure): deterministic, non-degenerate weights for parity tests.

The reference zero-initialises some convolutions (unet_openai.py:266-268, 324, 528), which would make a
fresh UNetModel output exactly 0; parity tests therefore fill every tensor from a per-name seed.
The same function is used by oracle/gen_golden.py (against the real reference) and by tests/.
"""
from __future__ import annotations

import zlib

import torch


def fill_state_dict(shapes: dict[str, tuple[int, ...]], gain: float = 1.0) -> dict[str, torch.Tensor]:
    out = {}
    for name, shape in shapes.items():
        g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()))
        shape = tuple(shape)
        if name.endswith(".bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # normalisation scale
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = gain * torch.randn(shape, generator=g) / fan_in ** 0.5
        out[name] = t.float()
    return out


def fill_autoencoder_state_dict(shapes: dict[str, tuple[int, ...]], gain: float = 1.6) -> dict[str, torch.Tensor]:
    """Weights for the BatchNorm autoencoders (h_vae_model_copy.py / h_vae_model.py) that keep the nets INPUT-SENSITIVE:
    running means small, running variances in [0.5, 1.5], and a weight gain that makes up for the LeakyReLU / ReLU
    attenuation, so the signal path carries the output (with `fill_state_dict`'s generic rule the running means land
    near 1, the ReLU nets die and the leaky ones attenuate the input to 1e-3 of the bias path -- a parity test on
    such weights checks little more than the biases)."""
    out = fill_state_dict(shapes, gain=gain)
    for name, shape in shapes.items():
        g = torch.Generator(device="cpu").manual_seed(zlib.crc32(name.encode()) ^ 0x5BD1E995)
        if name.endswith("running_mean"):
            out[name] = (0.1 * torch.randn(tuple(shape), generator=g)).float()
        elif name.endswith("running_var"):
            out[name] = (0.5 + torch.rand(tuple(shape), generator=g)).float()
    return out


def structured_images(b: int, ch: int, size: int, seed: int) -> torch.Tensor:
    """Deterministic test images in [0, 1] that differ from each other in structure and brightness (sinusoids,
    checkerboards, per-sample contrast) -- with i.i.d. noise images the pooled encoders see nearly the same input."""
    import math
    g = torch.Generator(device="cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, size), torch.linspace(0, 1, size), indexing="ij")
    out = []
    for i in range(b):
        f, ph = 1 + i % 4, 0.7 * i
        base = [torch.sin(2 * math.pi * f * xx + ph), torch.cos(2 * math.pi * f * yy - ph),
                ((xx * f * 2).floor() + (yy * f * 2).floor()) % 2][i % 3]
        img = torch.stack([base * (0.3 + 0.7 * ((i + c) % 3) / 2) for c in range(ch)])
        out.append(0.5 + 0.4 * img * (0.4 + 0.6 * (i % 5) / 4) + 0.05 * torch.randn(ch, size, size, generator=g))
    return torch.stack(out).clamp(0, 1).float()
