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
