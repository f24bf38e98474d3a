"""NumPy restatement of the Philox4x32-10 streams used by the CUDA path (TEST INFRASTRUCTURE).

These streams are OUR definition (the reference draws from Python's Mersenne Twister and the
third-party ``perlin_noise`` package); the oracle restates them independently so the device
implementation can be checked bit-for-bit (integers) / to rounding (derived normals).

Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
PURPOSE_OD, PURPOSE_PERLIN, PURPOSE_INTERP, PURPOSE_RESET = 1, 2, 3, 4


def philox4x32_10(key: int, c0, c1, c2, c3):
    """Vectorised over the counter words; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    return (x.astype(np.float64) + 1.0) * 2.3283064365386963e-10


def normal(key: int, c0, c1, c2, c3):
    x, y, _, _ = philox4x32_10(key, c0, c1, c2, c3)
    return np.sqrt(-2.0 * np.log(u01(x))) * np.cos(2 * np.pi * u01(y))


def od_noise(key: int, env, step, temp_std: float):
    """N(0, temp_std) keyed by (env, step) -- replaces ``random.gauss`` of environment.py:158."""
    return temp_std * normal(key, env, 0, step, PURPOSE_OD)


def perlin(key: int, env: int, x_over_period: float, nb_octaves: int, octaves_step: int) -> float:
    """1-D gradient noise with +-1 lattice gradients and the octave weights of perlin.py:41-56."""
    noise = 0.0
    for j in range(nb_octaves):
        xs = x_over_period * float((1 << j) * octaves_step)
        fl = np.floor(xs)
        f = xs - fl
        i0 = int(fl) & 0xFFFFFFFF
        a = philox4x32_10(key, env, i0, j, PURPOSE_PERLIN)[0]
        b = philox4x32_10(key, env, (i0 + 1) & 0xFFFFFFFF, j, PURPOSE_PERLIN)[0]
        g0 = 1.0 if int(a) & 1 else -1.0
        g1 = 1.0 if int(b) & 1 else -1.0
        fade = f * f * f * (f * (f * 6.0 - 15.0) + 10.0)
        v = g0 * f + fade * (g1 * (f - 1.0) - g0 * f)
        w = 1.0 / float(1 << j) if j < nb_octaves - 1 else 1.0 / float((1 << nb_octaves) - 1)
        noise += v * w
    return noise


def interp_choices(key: int, env: int, step: int, k: int, n: int):
    """``random.choices(ids, k)`` replacement keyed by (env, draw, step)."""
    x = philox4x32_10(key, env, np.arange(k), step, PURPOSE_INTERP)[0]
    return ((x.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
