"""Vectorised fp64 NumPy restatement of the reference environment step over [R, N].

TEST INFRASTRUCTURE (see ``oracle/__init__``).  Every function cites the reference lines it
follows (paths relative to ``/root/reference/server/app``).  House-level arithmetic keeps the
reference's literal operation order so agreement with the running reference is at the 1e-15
level; env-level scalars (outdoor temperature, solar gain, signal) are evaluated with Python
scalars exactly as the reference does.

State layout: per-house arrays are ``[R, N]`` (R independent replicas of an N-house cluster),
per-env arrays are ``[R]``.  All noise is injected by the caller.
"""
from __future__ import annotations

import datetime as _dt
import itertools
import math

import numpy as np

from .config import INTERP_GRIDS, INTERP_KEYS, INTERP_SHAPE, normalize_env_prop

EPOCH = _dt.datetime(1970, 1, 1)


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def to_epoch(dt: _dt.datetime) -> int:
    """Naive datetime -> integer seconds since 1970-01-01 (no time zone, like the reference)."""
    d = dt - EPOCH
    return d.days * 86400 + d.seconds


def from_epoch(sec: int) -> _dt.datetime:
    return EPOCH + _dt.timedelta(seconds=int(sec))


def deadband_l2(target, deadband, value):
    """utils/utils.py:4-23 (vectorised)."""
    hi = target + deadband / 2
    lo = target - deadband / 2
    return np.where(hi < value, (value - hi) ** 2, np.where(lo > value, (lo - value) ** 2, 0.0))


def solar_gain_scalar(dt: _dt.datetime, window_area: float, shading_coeff: float) -> float:
    """utils/utils.py:42-117 -- 18-term bivariate quartic, evaluated with Python floats."""
    x = dt.hour + dt.minute / 60 - 7.5
    if x < 0 or x > 10:
        scl = 0
    else:
        y = dt.month + dt.day / 30 - 1
        c = (
            4.36579418e01, 1.58055357e02, 8.76635241e01, -4.55944821e01, 3.24275366e00,
            -4.56096472e-01, -1.47795612e01, 4.68950855e00, -3.73313090e01, 5.78827663e00,
            1.04354810e00, 2.12969604e-02, 2.58881400e-03, -5.11397219e-04, 1.56398008e-02,
            -1.18302764e-01, -2.71446436e-01, -3.97855577e-02,
        )
        scl = (
            c[0] + x * c[1] + y * c[2] + x**2 * c[3] + x**2 * y * c[4] + x**2 * y**2 * c[5]
            + y**2 * c[6] + x * y**2 * c[7] + x * y * c[8] + x**3 * c[9] + y**3 * c[10]
            + x**3 * y * c[11] + x**3 * y**2 * c[12] + x**3 * y**3 * c[13]
            + x**2 * y**3 * c[14] + x * y**3 * c[15] + x**4 * c[16] + y**4 * c[17]
        )
    return window_area * shading_coeff * scl


def od_temp_scalar(dt: _dt.datetime, temp_prop: dict, noise: float) -> float:
    """core/environment/environment.py:132-159."""
    amplitude = (temp_prop["day_temp"] - temp_prop["night_temp"]) / 2.0
    bias = (temp_prop["day_temp"] + temp_prop["night_temp"]) / 2.0
    delay = -6.0 + temp_prop["phase"]
    time_day = dt.hour + dt.minute / 60.0
    temperature = amplitude * np.sin(2 * np.pi * (time_day + delay) / 24.0) + bias
    temperature += noise
    return float(temperature)


def hvac_fsm(on, lockout, sso, action, dt, dur):
    """core/environment/cluster/hvac.py:43-64 (vectorised lock-out state machine)."""
    on = on.astype(bool)
    action = action.astype(bool)
    sso = np.where(~on, sso + dt, sso)                       # :45-46
    lock = ~(on | (sso >= dur))                              # :48-51
    on_new = np.where(lock, False, action)                   # :53-56
    sso = np.where(~lock & on_new, 0, sso)                   # :57-58
    lock = lock | (~lock & ~on_new & (sso + dt < dur))       # :59-63
    return on_new, lock, sso


def thermal_update(t_air, t_mass, Ua, Ca, Cm, Hm, od_temp, Qa, dt):
    """core/environment/cluster/building.py:141-222, literal operation order, fp64."""
    od_K = od_temp + 273
    Ta_K = t_air + 273
    Tm_K = t_mass + 273
    Qm = 0
    a = Cm * Ca / Hm
    b = Cm * (Ua + Hm) / Hm + Ca
    c = Ua
    d = Qm + Qa + Ua * od_K
    g = Qm / Hm
    root = np.sqrt(b * b - 4 * a * c)
    r1 = (-b + root) / (2 * a)
    r2 = (-b - root) / (2 * a)
    dTA0dt = Hm * Tm_K / Ca - (Ua + Hm) * Ta_K / Ca + Ua * od_K / Ca + Qa / Ca
    A1 = (r2 * Ta_K - dTA0dt - r2 * d / c) / (r2 - r1)
    A2 = Ta_K - d / c - A1
    A3 = r1 * Ca / Hm + (Ua + Hm) / Hm
    A4 = r2 * Ca / Hm + (Ua + Hm) / Hm
    e1 = np.exp(r1 * dt)
    e2 = np.exp(r2 * dt)
    new_air = A1 * e1 + A2 * e2 + d / c
    new_mass = A1 * A3 * e1 + A2 * A4 * e2 + g + d / c
    return new_air - 273, new_mass - 273


# --------------------------------------------------------------------------------------
# neighbour tables (core/environment/cluster/agent_communication_builder.py)
# --------------------------------------------------------------------------------------
def nb_comm_of(n_agents: int, comm_prop: dict) -> int:
    """agent_communication_builder.py:49-52."""
    return int(min(comm_prop["max_nb_agents_communication"], n_agents - 1))


def comm_table(n_agents: int, comm_prop: dict, rng=None) -> np.ndarray:
    """Static neighbour table ``[N, nb_comm]`` for the non-per-step modes.

    ``neighbours`` :63-85, ``closed_groups`` :87-110, ``neighbours_2D`` :130-189,
    ``random_fixed`` :119-128 (needs ``rng`` = a ``random.Random``-like with ``sample``).
    ``random_sample`` has no static table (:112-117 returns ``{}``).
    """
    mode = comm_prop["mode"]
    c = nb_comm_of(n_agents, comm_prop)
    rows = []
    if mode == "neighbours":
        lo, hi = c // 2, (c + 1) // 2
        for i in range(n_agents):
            rows.append([(i - lo + j) % n_agents for j in range(lo)]
                        + [(i + 1 + j) % n_agents for j in range(hi)])
    elif mode == "closed_groups":
        for i in range(n_agents):
            base = i - (i % (c + 1))
            if base + c <= n_agents:
                ids = [base + j for j in range(comm_prop["max_nb_agents_communication"] + 1)]
            else:
                ids = [n_agents - c - 1 + j for j in range(c + 1)]
            ids.remove(i)
            rows.append(ids)
    elif mode == "neighbours_2D":
        row = comm_prop["row_size"]
        dist = comm_prop["max_communication_distance"]
        if n_agents % row != 0:
            raise ValueError("Neighbours 2D row_size must be a divisor of nb_agents")
        max_y = n_agents // row
        if dist >= (row + 1) // 2 or dist >= (max_y + 1) // 2:
            raise ValueError("Neighbours 2D distance_comm too large")
        pattern = [(dx, dy) for dx in range(-dist, dist + 1) for dy in range(-dist, dist + 1)
                   if abs(dx) + abs(dy) <= dist and (dx != 0 or dy != 0)]
        for i in range(n_agents):
            x, y = i % row, i // row
            rows.append([((y + dy) % max_y) * row + ((x + dx) % row) for dx, dy in pattern])
    elif mode == "random_fixed":
        for i in range(n_agents):
            ids = list(range(n_agents))
            ids.remove(i)
            rows.append(rng.sample(ids, k=c))
    elif mode == "random_sample":
        return np.zeros((n_agents, 0), dtype=np.int32)
    else:
        raise ValueError(f"unknown communication mode {mode}")
    width = {len(r) for r in rows}
    if len(width) != 1:
        raise ValueError("ragged neighbour table")
    return np.asarray(rows, dtype=np.int32).reshape(n_agents, -1)


# --------------------------------------------------------------------------------------
# interpolated base power (core/environment/power_grid/interpolation.py)
# --------------------------------------------------------------------------------------
def _nearest(grid, value):
    g = np.asarray(grid, dtype=np.float64)
    return np.argmin(np.abs(g - np.asarray(value)[..., None]), axis=-1)  # first min on ties


def interp_static_index(Ua, Cm, Ca, Hm, cap, default_house: dict) -> np.ndarray:
    """Nearest-neighbour part of ``interpolate_grid_fast`` (:137-167): the four thermal
    ratios (:227-231) and ``HVAC_power`` (:235,:160-162) after clipping (:245-264).
    Returns the flat index of the 5-D (air, mass, OD, hour, date) sub-table, row-major in
    key order ``(Ua, Cm, Ca, Hm, HVAC_power)``."""
    def clipped(v, key):
        g = INTERP_GRIDS[key]
        return np.clip(v, min(g), max(g))

    iu = _nearest(INTERP_GRIDS["Ua_ratio"], clipped(Ua / default_house["Ua"], "Ua_ratio"))
    icm = _nearest(INTERP_GRIDS["Cm_ratio"], clipped(Cm / default_house["Cm"], "Cm_ratio"))
    ica = _nearest(INTERP_GRIDS["Ca_ratio"], clipped(Ca / default_house["Ca"], "Ca_ratio"))
    ihm = _nearest(INTERP_GRIDS["Hm_ratio"], clipped(Hm / default_house["Hm"], "Hm_ratio"))
    ihv = _nearest(INTERP_GRIDS["HVAC_power"], clipped(cap, "HVAC_power"))
    return ((((iu * 3 + icm) * 3 + ica) * 3 + ihm) * 2 + ihv).astype(np.int64)


def _find_interval(grid, x):
    """scipy ``find_indices`` semantics: ``g[i] <= x < g[i+1]`` clipped to ``[0, n-2]``."""
    g = np.asarray(grid, dtype=np.float64)
    i = np.searchsorted(g, x, side="right") - 1
    i = np.clip(i, 0, g.size - 2)
    y = (x - g[i]) / (g[i + 1] - g[i])
    return i, y


def interp_sub_tables(table: np.ndarray) -> np.ndarray:
    """Re-order the 10-D table (key order ``INTERP_KEYS``) to ``[162, 9, 5, 8, 12, 6]``:
    axis 0 = (Ua, Cm, Ca, Hm, HVAC_power) nearest index, then the five linear dimensions."""
    t = np.asarray(table).reshape(INTERP_SHAPE)
    t = np.moveaxis(t, 7, 4)  # (Ua,Cm,Ca,Hm,HVAC, air,mass,OD,hour,date)
    return np.ascontiguousarray(t.reshape(162, 9, 5, 8, 12, 6))


def interp_point(sub: np.ndarray, static_idx, air, mass, od, hour, date):
    """5-D multilinear of ``scipy.interpolate.interpn`` (RegularGridInterpolator
    ``_evaluate_linear``): corners enumerated with the last dimension fastest, weight
    multiplied in dimension order, terms accumulated in enumeration order."""
    keys = ("air_temp", "mass_temp", "OD_temp", "hour", "date")
    xs = (air, mass, od, hour, date)
    idx, ys = [], []
    for k, x in zip(keys, xs):
        g = INTERP_GRIDS[k]
        x = np.clip(np.asarray(x, dtype=np.float64), min(g), max(g))  # :245-264
        i, y = _find_interval(g, x)
        idx.append(i)
        ys.append(y)
    value = np.zeros(np.broadcast(static_idx, *xs).shape)
    for corner in itertools.product((0, 1), repeat=5):
        weight = np.ones_like(value)
        for dim, up in enumerate(corner):
            weight = weight * (ys[dim] if up else (1 - ys[dim]))
        v = sub[static_idx, idx[0] + corner[0], idx[1] + corner[1], idx[2] + corner[2],
                idx[3] + corner[3], idx[4] + corner[4]]
        value = value + v * weight
    return value


# --------------------------------------------------------------------------------------
# controllers restated for closed-loop tests (core/agents/controllers)
# --------------------------------------------------------------------------------------
def deadband_bangbang(t_air, target, deadband, on):
    """bangbang_controllers.py:54-65 (and BasicController :75-89)."""
    return np.where(t_air < target - deadband / 2, False,
                    np.where(t_air > target + deadband / 2, True, on.astype(bool)))


def bangbang(t_air, target):
    """bangbang_controllers.py:75-82 (``BangBangController``)."""
    return t_air > target


def greedy_myopic(t_air, target, cap, cop, lockout, reg_signal):
    """greedy_myopic_controller.py:67-104 for one cluster (1-D inputs).

    Sort ascending by ``-(Ta - target)``; the reference uses pandas' default (unstable)
    quicksort, we use a stable sort -- results agree whenever the keys are distinct (Q13)."""
    key = -(t_air - target)
    order = np.argsort(key, kind="stable")
    power = cap / cop
    act = np.zeros(t_air.shape[0], dtype=bool)
    total = 0
    for h in order:
        p = power[h]
        if p + total < reg_signal or (abs(p + total - reg_signal) < abs(total - reg_signal)
                                      and not lockout[h]):
            total += p
            act[h] = True
    return act


# --------------------------------------------------------------------------------------
# the environment
# --------------------------------------------------------------------------------------
class NpOracle:
    """R replicas x N houses, fp64, all noise injected.

    ``state`` keys -- per house ``[R, N]``: ``t_air, t_mass, target, Ua, Ca, Cm, Hm, cap``
    (f64), ``on, lockout`` (bool), ``sso`` (int64); per env ``[R]``: ``epoch`` (int64 seconds),
    ``od_temp, signal, base_power, power, solar, artificial_ratio, max_power`` (f64),
    ``t_since_interp`` (int64).
    """

    def __init__(self, env_prop: dict | None, n_rep: int = 1, table: np.ndarray | None = None):
        self.p = normalize_env_prop(env_prop)
        self.R = int(n_rep)
        self.N = int(self.p["cluster_prop"]["nb_agents"])
        self.dt = self.p["time_step"]
        hp = self.p["cluster_prop"]["house_prop"]
        self.hp = hp
        self.hv = hp["hvac_prop"]
        self.nb_comm = nb_comm_of(self.N, self.p["cluster_prop"]["agents_comm_prop"])
        self.sub = interp_sub_tables(table) if table is not None else None
        self.state: dict = {}

    # ---- state ------------------------------------------------------------------------
    def set_state(self, st: dict) -> None:
        R, N = self.R, self.N
        f = lambda k: np.array(np.broadcast_to(np.asarray(st[k], dtype=np.float64), (R, N)))
        s = {k: f(k) for k in ("t_air", "t_mass", "target", "Ua", "Ca", "Cm", "Hm", "cap")}
        s["on"] = np.array(np.broadcast_to(np.asarray(st["on"]).astype(bool), (R, N)))
        s["lockout"] = np.array(np.broadcast_to(np.asarray(st["lockout"]).astype(bool), (R, N)))
        s["sso"] = np.array(np.broadcast_to(np.asarray(st["sso"]).astype(np.int64), (R, N)))
        e = lambda k, dflt=None: np.array(np.broadcast_to(
            np.asarray(st.get(k, dflt), dtype=np.float64), (R,)))
        s["epoch"] = np.array(np.broadcast_to(np.asarray(st["epoch"], dtype=np.int64), (R,)))
        s["od_temp"] = e("od_temp")
        s["signal"] = e("signal", 0.0)
        s["base_power"] = e("base_power", 0.0)
        s["artificial_ratio"] = e("artificial_ratio", self.p["power_grid_prop"]["artificial_ratio"])
        # cluster.py:63-65 -- cached from the *un-noised* hvac props (quirk Q2)
        s["max_power"] = e("max_power", self.N * (self.hv["cooling_capacity"] / self.hv["cop"]))
        period = self.p["power_grid_prop"]["base_power_props"]["interp_update_period"]
        s["t_since_interp"] = np.array(np.broadcast_to(
            np.asarray(st.get("t_since_interp", period + 1), dtype=np.int64), (R,)))
        s["solar"] = e("solar", 0.0)
        s["power"] = e("power", 0.0) if "power" in st else self.house_power(s).cumsum(axis=1)[:, -1]
        self.state = s

    def get_state(self) -> dict:
        return self.state

    def house_power(self, s=None):
        """hvac.py:101-111 + environment_properties.py:92-98."""
        s = self.state if s is None else s
        return np.where(s["on"], s["cap"] / self.hv["cop"], 0.0)

    # ---- one step ---------------------------------------------------------------------
    def step(self, actions, od_noise, perlin=None, interp_ids=None):
        """environment.py:72-108.  ``actions`` [R,N] truthy; ``od_noise`` [R] (the
        ``random.gauss`` draw of :158); ``perlin`` [R] (value of ``Perlin.calculate_noise``);
        ``interp_ids`` [R, k] house ids sampled at ``interpolation.py:223`` (only read on the
        steps where the interpolator fires and N > interp_nb_agents)."""
        s, p, R, N, dt = self.state, self.p, self.R, self.N, self.dt
        actions = np.broadcast_to(np.asarray(actions), (R, N))
        od_noise = np.broadcast_to(np.asarray(od_noise, dtype=np.float64), (R,))
        s["epoch"] = s["epoch"] + dt                                        # :87
        when = [from_epoch(e) for e in s["epoch"]]

        # Cluster.step (cluster.py:73-89): FSM, then thermal update with the PREVIOUS outdoor
        # temperature and the NEW datetime (quirk Q5)
        s["on"], s["lockout"], s["sso"] = hvac_fsm(
            s["on"], s["lockout"], s["sso"], actions, dt, self.hv["lockout_duration"])
        if self.hp["solar_gain"]:                                            # building.py:176-181
            s["solar"] = np.array([solar_gain_scalar(w, self.hp["window_area"],
                                                     self.hp["shading_coeff"]) for w in when])
        else:
            s["solar"] = np.zeros(R)
        q_hvac = np.where(s["on"], -1 * s["cap"] / (1 + self.hv["latent_cooling_fraction"]), 0)
        Qa = q_hvac + s["solar"][:, None]                                    # building.py:183-184
        s["t_air"], s["t_mass"] = thermal_update(
            s["t_air"], s["t_mass"], s["Ua"], s["Ca"], s["Cm"], s["Hm"],
            s["od_temp"][:, None], Qa, dt)
        house_p = self.house_power()
        s["power"] = house_p.cumsum(axis=1)[:, -1]                           # cluster.py:88 (in order)

        # outdoor temperature (environment.py:94)
        s["od_temp"] = np.array([od_temp_scalar(w, p["temp_prop"], n) for w, n in zip(when, od_noise)])

        # rewards with the OLD signal (environment.py:96-101, quirk Q6)
        rewards = self.rewards(s["power"], s["signal"])

        # power grid (environment.py:104-106)
        self.power_grid_step(when, perlin, interp_ids)
        return rewards

    # ---- rewards ----------------------------------------------------------------------
    def temp_penalty(self):
        """rewards_calculator.py:46-133."""
        s = self.state
        pp = self.p["reward_prop"]["penalty_props"]
        ind = deadband_l2(s["target"], self.hp["deadband"], s["t_air"])
        mode = pp["mode"]
        if mode == "individual_L2":
            return ind
        common = (ind / self.N).cumsum(axis=1)[:, -1:]                       # :60-66, in order
        cmax = np.maximum(ind.max(axis=1, keepdims=True), 0.0)               # :98-106
        if mode == "common_L2":
            return np.broadcast_to(common, ind.shape)
        if mode == "common_max_error":
            return np.broadcast_to(cmax, ind.shape)
        if mode == "mixture":                                                # :108-133
            a_i, a_c, a_m = pp["alpha_ind_l2"], pp["alpha_common_l2"], pp["alpha_common_max"]
            return (a_i * ind + a_c * common + a_m * cmax) / (a_i + a_c + a_m)
        raise ValueError(mode)

    def rewards(self, power, signal):
        """rewards_calculator.py:135-203."""
        rp = self.p["reward_prop"]
        sig_pen = ((power - signal) / self.N) ** 2                           # :198
        t0 = self.hp["target_temp"]
        norm_temp = float(deadband_l2(t0, 0, t0 + 1))                        # :155-159
        nrs = rp["norm_reg_sig"]
        norm_sig = float(deadband_l2(nrs, 0, 0.75 * nrs))                    # :161-165
        pen = self.temp_penalty()
        return -1 * (rp["alpha_temp"] * pen / norm_temp
                     + (rp["alpha_sig"] * sig_pen / norm_sig)[:, None])      # :174-179

    # ---- power grid -------------------------------------------------------------------
    def power_grid_step(self, when, perlin=None, interp_ids=None):
        """power_grid.py:80-102, :130-161; signal_calculator.py:33-129."""
        s, R, N = self.state, self.R, self.N
        gp = self.p["power_grid_prop"]
        bp, sp = gp["base_power_props"], gp["signal_properties"]
        if bp["mode"] == "constant":
            s["base_power"] = np.full(R, float(bp["avg_power_per_hvac"] * N))   # :145-148
        elif bp["mode"] == "interpolation":
            s["t_since_interp"] = s["t_since_interp"] + self.dt               # :150
            fire = s["t_since_interp"] >= bp["interp_update_period"]
            for r in np.nonzero(fire)[0]:
                ids = None if interp_ids is None else np.asarray(interp_ids)[r]
                s["base_power"][r] = self.interpolate_power(r, when[r], ids)
                s["t_since_interp"][r] = 0
        else:
            raise ValueError(bp["mode"])
        sig = np.zeros(R)
        for r in range(R):
            base, w = s["base_power"][r], when[r]
            t_sec = w.hour * 3600 + w.minute * 60 + w.second
            mode = sp["mode"]
            if mode == "flat":
                v = base
            elif mode == "sinusoidals":                                       # :46-76
                amps = [base * ratio for ratio in sp["amplitude_ratios"]]
                if len(sp["periods"]) != len(amps):
                    raise ValueError("periods and amplitude_ratios must have the same length")
                v = base
                for k, per in enumerate(sp["periods"]):
                    v += amps[k] * np.sin(2 * np.pi * t_sec / per)
            elif mode == "regular_steps":                                     # :78-98
                amplitude = sp["amplitude_per_hvac"] * N
                ratio = base / amplitude
                per = sp["period"]
                v = amplitude * np.heaviside((t_sec % per) - (1 - ratio) * per, 1)
            elif mode == "perlin":                                            # :100-115
                amp = sp["amplitude_ratios"][0]
                v = np.maximum(0, base + (base * amp * perlin[r]))
            else:
                raise ValueError(mode)
            v = v * s["artificial_ratio"][r]                                  # power_grid.py:99
            sig[r] = np.minimum(v, s["max_power"][r])                         # :100
        s["signal"] = sig

    def interpolate_power(self, r, when, ids):
        """interpolation.py:186-243 for replica ``r``."""
        s, N = self.state, self.N
        bp = self.p["power_grid_prop"]["base_power_props"]
        if self.hp["solar_gain"]:
            date = when.timetuple().tm_yday                                   # :206 (quirk Q12)
            hour = float(when.hour * 3600 + when.minute * 60 + when.second)   # :207-210
        else:
            date, hour = 0.0, 0.0
        k = bp["interp_nb_agents"]
        if N <= k:
            ids, factor = np.arange(N), 1.0                                   # :220-222
        else:
            ids, factor = np.asarray(ids, dtype=np.int64), float(N) / float(k)   # :223-225
        static = interp_static_index(s["Ua"][r, ids], s["Cm"][r, ids], s["Ca"][r, ids],
                                     s["Hm"][r, ids], s["cap"][r, ids], self.hp)
        vals = interp_point(
            self.sub, static,
            s["t_air"][r, ids] - s["target"][r, ids],
            s["t_mass"][r, ids] - s["target"][r, ids],
            s["od_temp"][r] - s["target"][r, ids],
            np.full(ids.shape, hour), np.full(ids.shape, float(date)))
        total = 0.0
        for v in vals:                                                        # :226-241 in order
            total += v
        return total * factor

    # ---- observations -----------------------------------------------------------------
    def obs_dim(self) -> int:
        sp, mp = self.p["state_prop"], self.p["cluster_prop"]["message_prop"]
        own = 10 + (2 if sp["hvac"] else 0) + (1 if sp["solar_gain"] else 0) + (5 if sp["thermal"] else 0)
        msg = 4 + (4 if mp["thermal"] else 0) + (3 if mp["hvac"] else 0)
        return own + msg * self.nb_comm

    def obs_vectors(self, table: np.ndarray | None = None) -> np.ndarray:
        """``norm_state_dict`` (utils/norm.py:178-218) applied to ``Environment.get_obs``
        (environment.py:110-130): returns ``[R, N, D]`` fp64.  ``table`` is the neighbour
        table ``[N, c]`` or per-replica ``[R, N, c]``; default = the static table of the
        configured mode."""
        s, R, N = self.state, self.R, self.N
        sp, mp = self.p["state_prop"], self.p["cluster_prop"]["message_prop"]
        nrs = self.p["reward_prop"]["norm_reg_sig"]
        dur = self.hv["lockout_duration"]
        hp = self.hp
        if table is None:
            table = comm_table(N, self.p["cluster_prop"]["agents_comm_prop"])
        table = np.broadcast_to(np.asarray(table), (R,) + tuple(np.asarray(table).shape[-2:]))
        cols = []
        rn = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), (R, N))
        # norm_hvac_dict :71-98
        cols += [rn(s["on"]), rn(s["lockout"]), rn((s["sso"] / dur).astype(np.int64)), rn(int(dur / dur))]
        if sp["hvac"]:
            cols += [rn(self.hv["cop"] / self.hv["cop"]),
                     rn(self.hv["latent_cooling_fraction"] / self.hv["latent_cooling_fraction"])]
        cols.append(rn((s["power"] / nrs)[:, None]))                          # norm_cluster_dict :139-148
        cols.append(rn((s["signal"] / (nrs * N))[:, None]))                   # norm_powergrid_dict :128-136
        # norm_building_dict :101-125
        cols += [rn(hp["deadband"]), (s["t_air"] - 20) / 5, (s["t_mass"] - 20) / 5, (s["target"] - 20) / 5]
        if sp["solar_gain"]:
            cols.append(rn((s["solar"] / 1000)[:, None]))
        if sp["thermal"]:
            cols += [s["Ua"] / hp["Ua"], s["Ca"] / hp["Ca"], s["Cm"] / hp["Cm"], s["Hm"] / hp["Hm"]]
            cols.append(rn(((s["od_temp"] - 20) / 5)[:, None]))              # env_norm_dict :164-165 (Q8)
        # norm_message :31-68 over Cluster.message (cluster.py:91-111) / Building.message (building.py:102-139)
        house_p = self.house_power()
        pmax = s["cap"] / self.hv["cop"]
        ridx = np.arange(R)[:, None]
        for k in range(table.shape[-1]):
            nb = table[:, :, k]
            g = lambda a: a[ridx, nb]
            cols.append(g(s["t_air"] - s["target"]) / 5)
            cols.append(g((s["sso"] / dur).astype(np.int64)).astype(np.float64))
            cols.append(g(house_p) / nrs)
            cols.append(g(pmax) / nrs)
            if mp["thermal"]:
                cols += [g(s["Ua"]) / hp["Ua"], g(s["Ca"]) / hp["Ca"], g(s["Cm"]) / hp["Cm"], g(s["Hm"]) / hp["Hm"]]
            if mp["hvac"]:                                                    # constants, quirk Q11
                cols += [rn(self.hv["cop"]), rn(self.hv["latent_cooling_fraction"]),
                         rn(self.hv["cooling_capacity"])]
        return np.stack([np.asarray(c, dtype=np.float64) for c in cols], axis=-1)
