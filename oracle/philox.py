"""Philox4x32-10 + Box-Muller restatement -- TEST INFRASTRUCTURE ONLY.

The reference draws its base samples with Julia's default RNG through Distributions.rand
(src/Flows.jl:167,181); that stream cannot be reproduced outside Julia, so the in-kernel generator
of `dflow_sample_rng` is specified HERE (counter-based Philox4x32-10, Salmon et al. 2011) and the
CUDA kernel is checked against this restatement:

  counter = (b_lo, b_hi, g, offset)   b = global sample index, g = group of 4 coordinates
  key     = (seed_lo, seed_hi)
  (r0,r1,r2,r3) = philox4x32_10(counter, key)
  u_i = ((r_i >> 8) + 0.5) * 2^-24                       in (0,1)
  z[4g+0] = sqrt(-2 ln u0) cos(2π u1) ; z[4g+1] = sqrt(-2 ln u0) sin(2π u1)
  z[4g+2] = sqrt(-2 ln u2) cos(2π u3) ; z[4g+3] = sqrt(-2 ln u2) sin(2π u3)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def normal_samples(d: int, B: int, seed: int, offset: int = 0, first_sample: int = 0) -> np.ndarray:
    """(d, B) float32 standard normals, sample b uses global index first_sample + b."""
    b = np.arange(first_sample, first_sample + B, dtype=np.uint64)
    b_lo = (b & MASK).astype(np.uint32)
    b_hi = (b >> np.uint64(32)).astype(np.uint32)
    out = np.empty((d, B), dtype=np.float32)
    for g in range((d + 3) // 4):
        r = philox4x32_10(b_lo, b_hi, np.uint32(g), np.uint32(offset), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        u = [((ri >> np.uint32(8)).astype(np.float64) + 0.5) * (2.0 ** -24) for ri in r]
        for p in range(2):
            rad = np.sqrt(-2.0 * np.log(u[2 * p]))
            ang = 2.0 * np.pi * u[2 * p + 1]
            for q, val in enumerate((rad * np.cos(ang), rad * np.sin(ang))):
                k = 4 * g + 2 * p + q
                if k < d:
                    out[k] = val.astype(np.float32)
    return out
