"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the device random-number
stream of gp_b200/csrc/rng.cu -- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11; the published round function and constants), 53-bit uniforms, Box-Muller.

The reference draws its normals with R's rnorm / MASS::mvrnorm (pendulum_fit.R:253) and
numpy.random.randn (ch2.py:43-45); neither stream is reproducible on a GPU, so parity for the
sampling row is: (a) the integer stream matches this restatement bit for bit,
(b) the normals match it to a few ulp, (c) draws equal mu + L z for that z, (d) moments.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, seed):
    """counter: uint64 array (the 64-bit counter in words c0, c1; c2 = c3 = 0); seed: 64-bit key.
    Returns an (len, 4) uint32 array."""
    ctr = np.asarray(counter, dtype=np.uint64)
    c0, c1 = ctr & MASK, ctr >> np.uint64(32)
    c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u53(a, b):
    m = ((a.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)) | (b.astype(np.uint64) >> np.uint64(6))
    return (m.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normals(seed, n, offset=0):
    """Normal numbers offset .. offset+n-1 of stream `seed` (offset even)."""
    assert offset % 2 == 0
    npairs = (n + 1) // 2
    r = philox4x32_10(np.arange(npairs, dtype=np.uint64) + np.uint64(offset // 2), seed)
    u1, u2 = u53(r[:, 0], r[:, 1]), u53(r[:, 2], r[:, 3])
    rad = np.sqrt(-2.0 * np.log(u1))
    out = np.empty(2 * npairs)
    out[0::2] = rad * np.cos(2.0 * np.pi * u2)
    out[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return out[:n]
