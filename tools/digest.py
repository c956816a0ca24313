"""Position-weighted 64-bit digests of int32 vectors, computable with numpy (fixtures,
CPU tests) and with torch on the device (GPU tests, bench.py) -- and additive over
shards, so ranks can digest their own part and sum.

    d_k(x) = sum_i x[i] * ((offset + i + 1) * C_k)   mod 2^64,   k = 0, 1

A full-size result (N = 2^26: 256 MiB) is checked against the committed digest of the
reference's own golden output (tests/golden/make_golden_large.py) without shipping it.
"""
from __future__ import annotations

import numpy as np

C = (0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F)  # odd 64-bit constants


def _signed(c: int) -> int:
    return c - (1 << 64) if c >= (1 << 63) else c


def digest_numpy(x: np.ndarray, offset: int = 0) -> tuple[int, int]:
    x = np.ascontiguousarray(x).reshape(-1)
    out = []
    with np.errstate(over="ignore"):
        idx = np.arange(offset + 1, offset + 1 + x.size, dtype=np.uint64)
        xv = x.astype(np.int64).astype(np.uint64)
        for c in C:
            out.append(int((xv * (idx * np.uint64(c))).sum(dtype=np.uint64)))
    return out[0], out[1]


def digest_torch(x, offset: int = 0):
    """Same digests for a torch int32 tensor (any device); returns a 2-element int64
    tensor holding the two's-complement images of d_0, d_1."""
    import torch
    x = x.reshape(-1)
    idx = torch.arange(offset + 1, offset + 1 + x.numel(), dtype=torch.int64, device=x.device)
    xv = x.to(torch.int64)
    return torch.stack([(xv * (idx * _signed(c))).sum() for c in C])


def as_unsigned(t) -> tuple[int, int]:
    return tuple(int(v) & ((1 << 64) - 1) for v in t.tolist())
