"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the counter-based generator the CUDA path
uses for its in-kernel draws, Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
algorithm torch's CUDA generator is built on as well), and of the dropout mask `sbm_dropout` derives from it
(csrc/sampler.cu `dropout_kernel`; replaces nn.Dropout at unet_openai.py:265).  Integer work: the CUDA kernel must match
bit for bit.  Pinned by the Random123 known-answer vectors (tests/test_oracle_cpu.py)."""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: tuple[int, int]) -> np.ndarray:
    """ctr: uint32 [..., 4]; key: two 32-bit words -> uint32 [..., 4]."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = key[0] & 0xFFFFFFFF, key[1] & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def words(seed: int, draw: int, quads: np.ndarray) -> np.ndarray:
    """The library's counter convention: counter = (quad lo, quad hi, draw lo, draw hi), key = seed -> uint32 [n, 4]."""
    q = quads.astype(np.uint64)
    ctr = np.stack([(q & MASK), (q >> np.uint64(32)), np.full_like(q, draw & 0xFFFFFFFF),
                    np.full_like(q, (draw >> 32) & 0xFFFFFFFF)], axis=-1).astype(np.uint32)
    return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def dropout_keep(seed: int, draw: int, rows: int, channels: int, p: float) -> np.ndarray:
    """keep mask [rows, channels] (bool) of `sbm_dropout`: element (row, c) uses 32-bit word number row * C8 + c of the
    stream (C8 = channels rounded up to 8), keep iff its top 24 bits >= floor(p * 2^24) in fp32."""
    c8 = (channels + 7) // 8 * 8
    n = rows * c8
    w = words(seed, draw, np.arange(n // 4, dtype=np.uint64)).reshape(rows, c8)
    thr = np.uint32(int(np.float32(p) * np.float32(16777216.0)))
    return ((w >> np.uint32(8)) >= thr)[:, :channels]
