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


PURPOSE_RESET_ENV = 5


def _tri(u, lo, hi, mode=1.0):
    """``random.triangular(lo, hi, mode)`` by CDF inversion (what the device kernel evaluates)."""
    fc = (mode - lo) / (hi - lo)
    return np.where(u < fc, lo + np.sqrt(u * (hi - lo) * (mode - lo)), hi - np.sqrt((1.0 - u) * (hi - lo) * (hi - mode)))


def reset_state(seed: int, n_rep: int, env_prop: dict, mode: str = "reference", quirk_ua: bool = True, rep_offset: int = 0):
    """NumPy restatement of ``k_reset`` (device-side Environment.reset): the
    distributions of building.py:224-267 / hvac.py:66-70 / environment.py:176-194 evaluated on the
    Philox streams keyed by (replica, house, draw).  Returns an oracle state dict."""
    from .config import normalize_env_prop
    from .np_oracle import from_epoch, od_temp_scalar, to_epoch

    p = normalize_env_prop(env_prop)
    hp, hv = p["cluster_prop"]["house_prop"], p["cluster_prop"]["house_prop"]["hvac_prop"]
    N, R, dt = p["cluster_prop"]["nb_agents"], n_rep, p["time_step"]
    env = (rep_offset + np.arange(R))[:, None]
    house = np.arange(N)[None, :]
    u0 = philox4x32_10(seed, env, house, 0, PURPOSE_RESET)
    u1 = philox4x32_10(seed, env, house, 1, PURPOSE_RESET)
    u2 = philox4x32_10(seed, env, house, 2, PURPOSE_RESET)
    h01 = lambda x: x.astype(np.float64) * 2.3283064365386963e-10
    g = np.sqrt(-2.0 * np.log(u01(u0[0]))) * np.cos(2 * np.pi * u01(u0[1]))
    npz = hp["noise_prop"]
    lo, hi = npz["factor_thermo_low"], npz["factor_thermo_high"]
    st = {"target": hp["target_temp"] + np.abs(npz["std_target_temp"] * g)}
    fu, fcm, fca, fhm = (_tri(h01(x), lo, hi) for x in (u0[2], u0[3], u1[0], u1[1]))
    st["Ua"] = fu if quirk_ua else hp["Ua"] * fu
    st["Cm"], st["Ca"], st["Hm"] = hp["Cm"] * fcm, hp["Ca"] * fca, hp["Hm"] * fhm
    caps = np.asarray(hv["noise_prop"]["cooling_capacity_list"], dtype=np.float64)
    ci = np.minimum(len(caps) - 1, ((u1[2].astype(np.uint64) * np.uint64(len(caps))) >> np.uint64(32)).astype(np.int64))
    st["cap"] = caps[ci]
    if mode == "reference":
        st["t_air"] = np.full((R, N), float(hp["init_air_temp"]))
        st["t_mass"] = np.full((R, N), float(hp["init_mass_temp"]))
        st["on"] = np.ones((R, N), dtype=bool)
        st["lockout"] = np.zeros((R, N), dtype=bool)
        st["sso"] = np.zeros((R, N), dtype=np.int64)
    else:
        st["t_air"] = st["target"] + (-2.0 + 6.0 * h01(u1[3]))
        st["t_mass"] = st["target"] + (-2.0 + 6.0 * h01(u2[0]))
        on = (u2[1] & np.uint32(1)).astype(bool)
        sso = np.where(on, 0, dt * (u2[2] & np.uint32(15)).astype(np.int64))
        st["on"], st["sso"] = on, sso
        st["lockout"] = (~on) & (sso < hv["lockout_duration"])
    e0 = philox4x32_10(seed, env[:, 0], 0, 0, PURPOSE_RESET_ENV)
    epoch = np.full(R, to_epoch(p["start_datetime"]), dtype=np.int64)
    if p["start_datetime_mode"] == "random":
        epoch = epoch + ((e0[0].astype(np.uint64) * np.uint64(364)) >> np.uint64(32)).astype(np.int64) * 86400 \
            + ((e0[1].astype(np.uint64) * np.uint64(86400)) >> np.uint64(32)).astype(np.int64)
    noise = p["temp_prop"]["temp_std"] * np.sqrt(-2.0 * np.log(u01(e0[2]))) * np.cos(2 * np.pi * u01(e0[3]))
    st["epoch"] = epoch
    st["od_temp"] = np.array([od_temp_scalar(from_epoch(e), p["temp_prop"], n) for e, n in zip(epoch, noise)])
    pmax0 = hv["cooling_capacity"] / hv["cop"]
    st["max_power"] = np.full(R, N * pmax0)
    st["power"] = np.full(R, N * pmax0) if mode == "reference" else np.where(st["on"], st["cap"] / hv["cop"], 0.0).sum(axis=1)
    st["signal"] = np.zeros(R)
    st["base_power"] = np.zeros(R)
    st["solar"] = np.zeros(R)
    return st
