"""Per-house pure-Python port of the reference environment step -- the CPU baseline ("port").

TEST INFRASTRUCTURE (see ``oracle/__init__``).  Unlike ``np_oracle`` this keeps the reference's
*cost model*: one record per house, a Python ``for`` over houses (cluster.py:82-88), per-house
solar-gain evaluation (building.py:176-181), the per-agent observation / message dict building
executed twice per step (cluster.py:89 + environment.py:108, quirk Q4) and the per-house reward
loop (rewards_calculator.py:171-180).  It is what ``bench.py`` times on the host cores as
``cpu_baseline`` / ``--impl reference`` (the reference itself cannot travel to the GPU box), and
it is pinned against the golden trajectories like the NumPy oracle.

Paths in citations are relative to ``/root/reference/server/app``.
"""
from __future__ import annotations

import datetime as _dt
import math

import numpy as np

from .config import normalize_env_prop
from .np_oracle import (comm_table, from_epoch, interp_point, interp_static_index, interp_sub_tables, nb_comm_of,
                        to_epoch)


def _deadband_l2(target, deadband, value):
    """utils/utils.py:4-23."""
    if target + deadband / 2 < value:
        return (value - (target + deadband / 2)) ** 2
    if target - deadband / 2 > value:
        return ((target - deadband / 2) - value) ** 2
    return 0.0


_SOLAR_C = (
    4.36579418e01, 1.58055357e02, 8.76635241e01, -4.55944821e01, 3.24275366e00, -4.56096472e-01,
    -1.47795612e01, 4.68950855e00, -3.73313090e01, 5.78827663e00, 1.04354810e00, 2.12969604e-02,
    2.58881400e-03, -5.11397219e-04, 1.56398008e-02, -1.18302764e-01, -2.71446436e-01, -3.97855577e-02,
)


def _solar_gain(when, window_area, shading_coeff):
    """utils/utils.py:42-117."""
    x = when.hour + when.minute / 60 - 7.5
    if x < 0 or x > 10:
        scl = 0
    else:
        y = when.month + when.day / 30 - 1
        c = _SOLAR_C
        scl = (c[0] + x * c[1] + y * c[2] + x**2 * c[3] + x**2 * y * c[4] + x**2 * y**2 * c[5] + y**2 * c[6]
               + x * y**2 * c[7] + x * y * c[8] + x**3 * c[9] + y**3 * c[10] + x**3 * y * c[11]
               + x**3 * y**2 * c[12] + x**3 * y**3 * c[13] + x**2 * y**3 * c[14] + x * y**3 * c[15]
               + x**4 * c[16] + y**4 * c[17])
    return window_area * shading_coeff * scl


class ScalarEnv:
    """One cluster (R = 1), houses as dict records, all noise injected."""

    def __init__(self, env_prop, table=None):
        self.p = normalize_env_prop(env_prop)
        self.N = int(self.p["cluster_prop"]["nb_agents"])
        self.dt = self.p["time_step"]
        self.hp = self.p["cluster_prop"]["house_prop"]
        self.hv = self.hp["hvac_prop"]
        self.comm = comm_table(self.N, self.p["cluster_prop"]["agents_comm_prop"]) \
            if self.p["cluster_prop"]["agents_comm_prop"]["mode"] not in ("random_fixed", "random_sample") else None
        self.sub = interp_sub_tables(table) if table is not None else None
        self.houses = []

    # ---- state ------------------------------------------------------------------------
    def set_state(self, st):
        g = lambda k, i: np.asarray(st[k]).reshape(-1)[i]
        self.houses = [dict(
            t_air=float(g("t_air", i)), t_mass=float(g("t_mass", i)), target=float(g("target", i)),
            Ua=float(g("Ua", i)), Ca=float(g("Ca", i)), Cm=float(g("Cm", i)), Hm=float(g("Hm", i)),
            cap=float(g("cap", i)), on=bool(g("on", i)), lockout=bool(g("lockout", i)), sso=int(g("sso", i)),
            solar=float(np.asarray(st.get("solar", 0.0)).reshape(-1)[0]),
        ) for i in range(self.N)]
        e = lambda k, d=0.0: float(np.asarray(st.get(k, d)).reshape(-1)[0])
        self.when = from_epoch(int(np.asarray(st["epoch"]).reshape(-1)[0]))
        self.od_temp = e("od_temp")
        self.signal = e("signal")
        self.base_power = e("base_power")
        self.artificial_ratio = e("artificial_ratio", self.p["power_grid_prop"]["artificial_ratio"])
        self.max_power = e("max_power", self.N * (self.hv["cooling_capacity"] / self.hv["cop"]))
        self.power = e("power") if "power" in st else sum(self._house_power(h) for h in self.houses)
        period = self.p["power_grid_prop"]["base_power_props"]["interp_update_period"]
        self.t_since_interp = int(np.asarray(st.get("t_since_interp", period + 1)).reshape(-1)[0])

    reinject = set_state

    def get_state(self):
        arr = lambda k, dt=np.float64: np.array([[h[k] for h in self.houses]], dtype=dt)
        return dict(t_air=arr("t_air"), t_mass=arr("t_mass"), on=arr("on", bool), lockout=arr("lockout", bool),
                    sso=arr("sso", np.int64), epoch=np.array([to_epoch(self.when)]), od_temp=np.array([self.od_temp]),
                    signal=np.array([self.signal]), base_power=np.array([self.base_power]),
                    power=np.array([self.power]), solar=np.array([self.houses[0]["solar"]]))

    def _house_power(self, h):
        """hvac.py:101-111."""
        return h["cap"] / self.hv["cop"] if h["on"] else 0.0

    # ---- one step (environment.py:72-108) ---------------------------------------------
    def step(self, actions, od_noise, perlin=None, interp_ids=None, comm=None):
        actions = np.asarray(actions).reshape(-1)
        dt, dur = self.dt, self.hv["lockout_duration"]
        self.when = self.when + _dt.timedelta(seconds=dt)
        # Cluster.step (cluster.py:73-89)
        self.power = 0.0
        for i, h in enumerate(self.houses):
            action = bool(actions[i])
            # HVAC.step (hvac.py:43-64)
            if not h["on"]:
                h["sso"] += dt
            if h["on"] or h["sso"] >= dur:
                h["lockout"] = False
            else:
                h["lockout"] = True
            if h["lockout"]:
                h["on"] = False
            else:
                h["on"] = action
                if h["on"]:
                    h["sso"] = 0
                elif h["sso"] + dt < dur:
                    h["lockout"] = True
            self._update_temperature(h)
            self.power += self._house_power(h)
        self._obs_dicts(comm)                       # discarded (cluster.py:89)
        # compute_od_temp (environment.py:132-159)
        tp = self.p["temp_prop"]
        amplitude = (tp["day_temp"] - tp["night_temp"]) / 2.0
        bias = (tp["day_temp"] + tp["night_temp"]) / 2.0
        time_day = self.when.hour + self.when.minute / 60.0
        temperature = amplitude * np.sin(2 * np.pi * (time_day + (-6.0 + tp["phase"])) / 24.0) + bias
        temperature += float(np.asarray(od_noise).reshape(-1)[0])
        self.od_temp = float(temperature)
        rewards = self._rewards()
        self._power_grid_step(None if perlin is None else float(np.asarray(perlin).reshape(-1)[0]), interp_ids)
        self.last_obs = self._obs_dicts(comm)
        return np.array([[rewards[i] for i in range(self.N)]])

    def _update_temperature(self, h):
        """building.py:141-222."""
        Hm, Ca, Ua, Cm = h["Hm"], h["Ca"], h["Ua"], h["Cm"]
        od_K, ta_K, tm_K = self.od_temp + 273, h["t_air"] + 273, h["t_mass"] + 273
        q = -1 * h["cap"] / (1 + self.hv["latent_cooling_fraction"]) if h["on"] else 0
        h["solar"] = _solar_gain(self.when, self.hp["window_area"], self.hp["shading_coeff"]) \
            if self.hp["solar_gain"] else 0.0
        Qa = q + h["solar"]
        Qm = 0
        a = Cm * Ca / Hm
        b = Cm * (Ua + Hm) / Hm + Ca
        c = Ua
        d = Qm + Qa + Ua * od_K
        g = Qm / Hm
        r1 = (-b + np.sqrt(b**2 - 4 * a * c)) / (2 * a)
        r2 = (-b - np.sqrt(b**2 - 4 * a * c)) / (2 * a)
        dTA0dt = Hm * tm_K / Ca - (Ua + Hm) * ta_K / Ca + Ua * od_K / Ca + Qa / Ca
        A1 = (r2 * ta_K - dTA0dt - r2 * d / c) / (r2 - r1)
        A2 = ta_K - d / c - A1
        A3 = r1 * Ca / Hm + (Ua + Hm) / Hm
        A4 = r2 * Ca / Hm + (Ua + Hm) / Hm
        h["t_air"] = A1 * np.exp(r1 * self.dt) + A2 * np.exp(r2 * self.dt) + d / c - 273
        h["t_mass"] = A1 * A3 * np.exp(r1 * self.dt) + A2 * A4 * np.exp(r2 * self.dt) + g + d / c - 273

    def _rewards(self):
        """rewards_calculator.py:135-203."""
        rp, N = self.p["reward_prop"], self.N
        pp = rp["penalty_props"]
        sig_pen = ((self.power - self.signal) / N) ** 2
        t0 = self.hp["target_temp"]
        norm_temp = _deadband_l2(t0, 0, t0 + 1)
        norm_sig = _deadband_l2(rp["norm_reg_sig"], 0, 0.75 * rp["norm_reg_sig"])
        db = self.hp["deadband"]
        out = {}
        for i, h in enumerate(self.houses):
            mode = pp["mode"]
            ind = _deadband_l2(h["target"], db, h["t_air"])
            if mode == "individual_L2":
                pen = ind
            else:
                common = 0.0
                cmax = 0.0
                for o in self.houses:                      # O(N^2) overall, as in the reference
                    v = _deadband_l2(o["target"], db, o["t_air"])
                    common += v / N
                    if v > cmax:
                        cmax = v
                if mode == "common_L2":
                    pen = common
                elif mode == "common_max_error":
                    pen = cmax
                else:
                    a_i, a_c, a_m = pp["alpha_ind_l2"], pp["alpha_common_l2"], pp["alpha_common_max"]
                    pen = (a_i * ind + a_c * common + a_m * cmax) / (a_i + a_c + a_m)
            out[i] = -1 * (rp["alpha_temp"] * pen / norm_temp + rp["alpha_sig"] * sig_pen / norm_sig)
        return out

    def _power_grid_step(self, perlin, interp_ids):
        """power_grid.py:80-102,130-161 + signal_calculator.py:33-129."""
        gp = self.p["power_grid_prop"]
        bp, sp = gp["base_power_props"], gp["signal_properties"]
        N, w = self.N, self.when
        if bp["mode"] == "constant":
            self.base_power = bp["avg_power_per_hvac"] * N
        else:
            self.t_since_interp += self.dt
            if self.t_since_interp >= bp["interp_update_period"]:
                self.base_power = self._interpolate(interp_ids)
                self.t_since_interp = 0
        base = self.base_power
        t_sec = w.hour * 3600 + w.minute * 60 + w.second
        mode = sp["mode"]
        if mode == "flat":
            v = base
        elif mode == "sinusoidals":
            v = base
            for k, per in enumerate(sp["periods"]):
                v += base * sp["amplitude_ratios"][k] * np.sin(2 * np.pi * t_sec / per)
        elif mode == "regular_steps":
            amplitude = sp["amplitude_per_hvac"] * N
            v = amplitude * np.heaviside((t_sec % sp["period"]) - (1 - base / amplitude) * sp["period"], 1)
        else:
            v = np.maximum(0, base + (base * sp["amplitude_ratios"][0] * perlin))
        v = v * self.artificial_ratio
        self.signal = float(np.minimum(v, self.max_power))

    def _interpolate(self, ids):
        """interpolation.py:186-243."""
        bp = self.p["power_grid_prop"]["base_power_props"]
        N, k = self.N, bp["interp_nb_agents"]
        if self.hp["solar_gain"]:
            date = self.when.timetuple().tm_yday
            hour = float(self.when.hour * 3600 + self.when.minute * 60 + self.when.second)
        else:
            date, hour = 0.0, 0.0
        if N <= k:
            ids, factor = list(range(N)), 1.0
        else:
            ids, factor = [int(i) for i in np.asarray(ids).reshape(-1)], float(N) / float(k)
        total = 0.0
        for i in ids:
            h = self.houses[i]
            s = interp_static_index(np.array(h["Ua"]), np.array(h["Cm"]), np.array(h["Ca"]), np.array(h["Hm"]),
                                    np.array(h["cap"]), self.hp)
            total += float(interp_point(self.sub, s, h["t_air"] - h["target"], h["t_mass"] - h["target"],
                                        self.od_temp - h["target"], hour, float(date)))
        return total * factor

    # ---- observation dicts (cluster.py:91-121, building.py:79-139, environment.py:110-130) ---
    def _obs_dicts(self, comm=None):
        table = self.comm if comm is None else comm
        hv = self.hv
        out = {}
        for i, h in enumerate(self.houses):
            msgs = []
            for j in (table[i] if table is not None else ()):
                o = self.houses[int(j)]
                msgs.append({
                    "seconds_since_off": o["sso"],
                    "curr_consumption": self._house_power(o),
                    "max_consumption": o["cap"] / hv["cop"],
                    "lockout_duration": hv["lockout_duration"],
                    "current_temp_diff_to_target": o["t_air"] - o["target"],
                })
            out[i] = {
                "turned_on": h["on"], "seconds_since_off": h["sso"], "lockout": h["lockout"], "cop": hv["cop"],
                "cooling_capacity": h["cap"], "latent_cooling_fraction": hv["latent_cooling_fraction"],
                "lockout_duration": hv["lockout_duration"], "target_temp": h["target"],
                "deadband": self.hp["deadband"], "Ua": h["Ua"], "Ca": h["Ca"], "Cm": h["Cm"], "Hm": h["Hm"],
                "indoor_temp": h["t_air"], "mass_temp": h["t_mass"], "solar_gain": h["solar"],
                "cluster_hvac_power": self.power, "message": msgs, "OD_temp": self.od_temp,
                "datetime": self.when, "reg_signal": self.signal,
            }
        return out
