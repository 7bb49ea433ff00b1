"""Counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) and
the normal stream the library's noise-augmentation kernel draws from it.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference adds ``0.1 * torch.randn_like(naction)`` to the actions
(il-and-cil/torch_nfbc_ant.py:264-265); torch's CUDA generator is Philox4x32-10 as well, but its stream layout is an
implementation detail of ATen, so parity for this step is: the generator itself bit-exact against the published
known-answer vectors (Random123 kat_vectors), the kernel's normals equal to this restatement, and the moments right.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32 counters, key: (2,) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def u01(r):
    """24 random bits -> a float in (0, 1), as the kernel maps them."""
    return (r >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0) + 0.5 / 16777216.0


def normals(n, seed, offset=0):
    """The n standard normals of nf_noise_augment(seed, offset): group g = i // 4 uses counter offset + g."""
    groups = (n + 3) // 4
    g = np.arange(groups, dtype=np.uint64) + np.uint64(offset)
    ctr = np.stack([g & MASK, g >> np.uint64(32), np.zeros_like(g), np.zeros_like(g)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    r0, r1 = np.sqrt(-2.0 * np.log(u01(r[:, 0]))), np.sqrt(-2.0 * np.log(u01(r[:, 2])))
    a0, a1 = 2.0 * np.pi * u01(r[:, 1]), 2.0 * np.pi * u01(r[:, 3])
    z = np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1)
    return z.reshape(-1)[:n]
