"""Specification of the device shuffle (dflow_shuffle_indices) -- TEST INFRASTRUCTURE ONLY.

The reference shuffles with Random.randperm (DataPartition, src/Data.jl:112-128) and Flux.DataLoader(shuffle=true)
(src/Flows.jl:394); which permutation is drawn is an implementation detail of Julia's RNG, so the drop-in only has to draw
*a* uniform-looking permutation reproducibly.  The CUDA path uses a stateless bijection so that every rank can evaluate
its own slice: a 6-round balanced Feistel network over 2*hb bits (hb = ceil(ceil(log2 n) / 2)) with a murmur3-finaliser
round function, keys from splitmix64(seed), cycle-walked into [0, n).  This file restates it in NumPy integer arithmetic;
tests compare bit-exactly.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _mix(x):
    x = x.astype(np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & M32
    x ^= x >> np.uint64(16)
    return x


def keys(seed: int):
    z = seed & 0xFFFFFFFFFFFFFFFF
    out = []
    for _ in range(6):
        z = (z + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        y = z
        y = ((y ^ (y >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        y = ((y ^ (y >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        y ^= y >> 31
        out.append(y >> 32)
    return out


def half_bits(n: int) -> int:
    bits = 1
    while bits < 62 and (1 << bits) < n:
        bits += 1
    return max(1, (bits + 1) // 2)


def permutation(seed: int, n: int, first: int = 0, count: int = None) -> np.ndarray:
    """perm(first), ..., perm(first + count - 1) of the seed's permutation of [0, n)."""
    count = n - first if count is None else count
    hb = half_bits(n)
    mask = np.uint64((1 << hb) - 1)
    ks = [np.uint64(k) for k in keys(seed)]
    x = np.arange(first, first + count, dtype=np.uint64)
    todo = np.ones(count, bool)
    while todo.any():
        v = x[todo]
        l, r = (v >> np.uint64(hb)) & mask, v & mask
        for q in range(6):
            t = l ^ (_mix(r ^ ks[q]) & mask)
            l, r = r, t
        v = (l << np.uint64(hb)) | r
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)
