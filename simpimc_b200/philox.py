"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) in numpy -- the host mirror of the random
stream the device-resident moves draw from (csrc/mc.cuh).  counter = (attempt_lo, attempt_hi,
clone, slot), key = (seed_lo, seed_hi)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays; returns four uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def uniform_from_bits(a, b):
    """Uniform double in (0, 1] from two uint32 words, as UniformFromBits in csrc/mc.cuh."""
    x = (np.asarray(a, dtype=np.uint64) >> np.uint64(5)) << np.uint64(26) | (np.asarray(b, dtype=np.uint64) >> np.uint64(6))
    return (x.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
