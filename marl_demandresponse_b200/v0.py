"""Adapter for the reference's legacy gym-style environment (SURVEY 8b "v0 adapter", 8f-4).

``MADemandResponseEnv(config, test=False)`` mirrors ``server/v0/env/MA_DemandResponse.py:23-180``:
``reset() -> obs_dict`` and ``step(action_dict) -> (obs_dict, rewards_dict, dones_dict, info_dict)``
with the legacy observation keys (``house_temp``, ``hvac_turned_on``, ... :888-917), the legacy
message dicts (:592-614), ``info_dict = {"cluster_hvac_power": P}`` and all-``False`` dones (:342-358),
driven by the legacy ``config_dict`` (``server/v0/config.py``).  The dynamics are the same GPU step as
the drop-in :class:`~marl_demandresponse_b200.environment.Environment` -- the legacy thermal model,
lock-out state machine, power grid and reward are the ones the app inherited -- what differs, and is
reproduced here, is everything around the step:

* construction / ``reset()`` both run ``build_environment`` (:84-119), i.e. consume Python's
  ``random`` twice, in the legacy order: per house two start-temperature draws, the target draw, four
  triangular factors and the capacity choice (``v0/utils.py:431-482``), then the start date (:509-518),
  per HVAC the lock-out noise ``randint`` (:396-398), the phase, the first outdoor temperature (:994-1018),
  ``random_fixed`` neighbour samples (:803-808), the artificial-ratio draw and the perlin seed (:1052-1125),
  and the sampled houses of the first interpolation (:1153-1154);
* the HVACs start OFF with ``seconds_since_off = lockout_duration`` (:401-403, quirk Q3), the start
  temperatures ARE noised, ``Ua`` is multiplied by its factor (``v0/utils.py:453``) and the cluster's
  ``max_power`` is the sum of the NOISED capacities (:752-757);
* ``random_sample`` neighbours are drawn once per step (:918-932), after the outdoor-temperature draw;
* ``normStateDict`` (``v0/utils.py:541-657``) keeps ``seconds_since_off / lockout_duration`` as a float and
  divides the cluster power by ``norm_reg_sig * nb_agents`` (quirks Q9, Q10): :func:`norm_state_dict_v0`.

``random.seed(s)`` before construction therefore reproduces the legacy trajectory
(``tests/test_gpu_v0.py`` against ``tests/golden/v0_*.npz`` recorded from the reference).  Per-HVAC
``lockout_noise != 0`` gives every HVAC its own lock-out duration (:397-403): the durations travel to the device as
a per-house plane and the step then runs on the general path (the fused tile kernels share one duration).
"""
from __future__ import annotations

import copy
import datetime as _dt
import random
from typing import Any, Dict, List

import numpy as np

from .core import DrSim, comm_width, flatten_config, to_epoch
from .environment import Environment, _Perlin, build_comm_table
from .properties import EnvironmentProperties


def _by_key(d: Dict[Any, Any], key):
    """Dict lookup that also accepts the string keys a JSON round trip leaves behind."""
    return d[key] if key in d else d[str(key)]


def props_from_v0(config: Dict[str, Any], test: bool = False) -> EnvironmentProperties:
    """The legacy ``config_dict`` expressed as the app-style property tree the C ABI is configured from."""
    env, house, hvac = config["default_env_prop"], config["default_house_prop"], config["default_hvac_prop"]
    nh = config["noise_house_prop_test" if test else "noise_house_prop"]
    nv = config["noise_hvac_prop_test" if test else "noise_hvac_prop"]
    nhp = nh["noise_parameters"][nh["noise_mode"]]
    nvp = nv["noise_parameters"][nv["noise_mode"]]
    cl, pg, rw = env["cluster_prop"], env["power_grid_prop"], env["reward_prop"]
    temp = cl["temp_parameters"][cl["temp_mode"]]
    sig_mode = pg["signal_mode"]
    sp = pg["signal_parameters"][sig_mode]
    signal: Dict[str, Any] = {"mode": "perlin" if "perlin" in sig_mode else sig_mode}
    if "perlin" in sig_mode:        # scalar amplitude in the legacy config (quirk Q15)
        signal.update(amplitude_ratios=[float(sp["amplitude_ratios"])], nb_octaves=sp["nb_octaves"],
                      octaves_step=sp["octaves_step"], period=sp["period"])
    elif sig_mode == "sinusoidals":
        signal.update(amplitude_ratios=list(sp["amplitude_ratios"]), periods=list(sp["periods"]))
    elif sig_mode == "regular_steps":
        signal.update(amplitude_per_hvac=sp["amplitude_per_hvac"], period=sp["period"])
    pen_mode = rw["temp_penalty_mode"]
    mix = rw["temp_penalty_parameters"].get("mixture", {})
    base_c, base_i = pg["base_power_parameters"]["constant"], pg["base_power_parameters"]["interpolation"]
    comm2d = cl.get("agents_comm_parameters", {}).get("neighbours_2D", {})
    start = env["start_datetime"]
    if isinstance(start, str):
        start = _dt.datetime.strptime(start, "%Y-%m-%d %H:%M:%S")
    return EnvironmentProperties(**{
        "start_datetime": start, "start_datetime_mode": env["start_datetime_mode"],
        "time_step": _dt.timedelta(seconds=env["time_step"]),
        "temp_prop": {"day_temp": temp["day_temp"], "night_temp": temp["night_temp"], "temp_std": temp["temp_std"],
                      "random_phase_offset": bool(temp.get("random_phase_offset", False)), "phase": 0.0},
        "state_prop": dict(env["state_properties"]),
        "reward_prop": {"alpha_temp": rw["alpha_temp"], "alpha_sig": rw["alpha_sig"], "norm_reg_sig": rw["norm_reg_sig"],
                        "penalty_props": {"mode": "common_max_error" if pen_mode.startswith("common_max") else pen_mode,
                                          "alpha_ind_l2": mix.get("alpha_ind_L2", 1.0),
                                          "alpha_common_l2": mix.get("alpha_common_L2", 1.0),
                                          "alpha_common_max": mix.get("alpha_common_max", 0.0)}},
        "cluster_prop": {
            "nb_agents": cl["nb_agents"], "nb_agents_comm": cl["nb_agents_comm"],
            "agents_comm_prop": {"mode": cl["agents_comm_mode"], "max_nb_agents_communication": cl["nb_agents_comm"],
                                 "row_size": comm2d.get("row_size", 5),
                                 "max_communication_distance": comm2d.get("distance_comm", 2)},
            "message_prop": dict(env["message_properties"]),
            "house_prop": {
                "Ua": house["Ua"], "Ca": house["Ca"], "Hm": house["Hm"], "Cm": house["Cm"],
                "target_temp": house["target_temp"], "deadband": house["deadband"],
                "init_air_temp": house["init_air_temp"], "init_mass_temp": house["init_mass_temp"],
                "solar_gain": bool(house["solar_gain_bool"]), "window_area": house["window_area"],
                "shading_coeff": house["shading_coeff"],
                "noise_prop": {k: nhp[k] for k in ("std_start_temp", "std_target_temp", "factor_thermo_low", "factor_thermo_high")},
                "hvac_prop": {"cop": hvac["COP"], "cooling_capacity": hvac["cooling_capacity"],
                              "latent_cooling_fraction": hvac["latent_cooling_fraction"],
                              "lockout_duration": hvac["lockout_duration"],
                              "noise_prop": {"lockout_noise": hvac.get("lockout_noise", 0),
                                             "cooling_capacity_list": list(_by_key(nvp["cooling_capacity_list"], hvac["cooling_capacity"]))}},
            }},
        "power_grid_prop": {
            "artificial_signal_ratio_range": pg["artificial_signal_ratio_range"], "artificial_ratio": pg["artificial_ratio"],
            "base_power_props": {"mode": pg["base_power_mode"], "avg_power_per_hvac": base_c["avg_power_per_hvac"],
                                 "init_signal_per_hvac": base_c["init_signal_per_hvac"],
                                 "path_datafile": base_i["path_datafile"],
                                 "interp_update_period": base_i["interp_update_period"],
                                 "interp_nb_agents": base_i["interp_nb_agents"]},
            "signal_properties": signal},
    })


def norm_state_dict_v0(s: Dict[str, Any], config: Dict[str, Any], return_dict: bool = False):
    """``v0/utils.py:541-657`` for one agent's legacy observation dict."""
    house, hvac, env = config["default_house_prop"], config["default_hvac_prop"], config["default_env_prop"]
    st, mp = env["state_properties"], env["message_properties"]
    nrs, n = env["reward_prop"]["norm_reg_sig"], env["cluster_prop"]["nb_agents"]
    out: Dict[str, Any] = {}
    temps = ["house_temp", "house_mass_temp", "house_target_temp"] + (["OD_temp"] if st["thermal"] else [])
    for k in temps:
        out[k] = (s[k] - 20) / 5
    out["house_deadband"] = s["house_deadband"]
    if st["day"]:
        day = s["datetime"].timetuple().tm_yday
        out["sin_day"], out["cos_day"] = np.sin(day * 2 * np.pi / 365), np.cos(day * 2 * np.pi / 365)
    if st["hour"]:
        hour = s["datetime"].hour
        out["sin_hr"], out["cos_hr"] = np.sin(hour * 2 * np.pi / 24), np.cos(hour * 2 * np.pi / 24)
    if st["solar_gain"]:
        out["house_solar_gain"] = s["house_solar_gain"] / 1000
    ratios = ["hvac_cooling_capacity"]
    if st["thermal"]:
        ratios += ["house_Ua", "house_Cm", "house_Ca", "house_Hm"]
    if st["hvac"]:
        ratios += ["hvac_COP", "hvac_latent_cooling_fraction"]
    for k in ratios:
        name = k.split("_", 1)[1]
        out[k] = s[k] / (house[name] if name in house else hvac[name])
    out["hvac_turned_on"] = 1 if s["hvac_turned_on"] else 0
    out["hvac_lockout"] = 1 if s["hvac_lockout"] else 0
    out["hvac_seconds_since_off"] = s["hvac_seconds_since_off"] / s["hvac_lockout_duration"]      # float (quirk Q9)
    out["hvac_lockout_duration"] = s["hvac_lockout_duration"] / s["hvac_lockout_duration"]
    out["reg_signal"] = s["reg_signal"] / (nrs * n)
    out["cluster_hvac_power"] = s["cluster_hvac_power"] / (nrs * n)                                # quirk Q10
    msgs = []
    for m in s["message"]:
        r = {"current_temp_diff_to_target": m["current_temp_diff_to_target"] / 5,
             "hvac_seconds_since_off": m["hvac_seconds_since_off"] / s["hvac_lockout_duration"],
             "hvac_curr_consumption": m["hvac_curr_consumption"] / nrs,
             "hvac_max_consumption": m["hvac_max_consumption"] / nrs}
        if mp["thermal"]:
            for k in ("Ua", "Cm", "Ca", "Hm"):
                r["house_" + k] = m["house_" + k] / house[k]
        if mp["hvac"]:
            r["hvac_COP"] = m["hvac_COP"] / hvac["COP"]
            r["hvac_latent_cooling_fraction"] = m["hvac_latent_cooling_fraction"] / hvac["latent_cooling_fraction"]
            r["hvac_cooling_capacity"] = m["hvac_cooling_capacity"] / hvac["cooling_capacity"]
        msgs.append(r)
    if return_dict:
        out["message"] = msgs
        return out
    flat: List[float] = []
    for r in msgs:
        flat += list(r.values())
    return np.array(list(out.values()) + flat)


class MADemandResponseEnv(Environment):
    """GPU-backed stand-in for the legacy ``MADemandResponseEnv`` (``MA_DemandResponse.py:23``)."""

    def __init__(self, config: Dict[str, Any], test: bool = False, device: int = 0, precision: str = "f64",
                 interp_table=None) -> None:
        self.config = config
        self.test = test
        self.init_props = props_from_v0(config, test)
        self.n = int(self.init_props.cluster_prop.nb_agents)
        self.agent_ids = list(range(self.n))
        self.nb_agents = self.n
        self._device, self._precision = device, precision
        self._interp_table = interp_table
        if self.init_props.power_grid_prop.base_power_props.mode == "interpolation" and interp_table is None:
            self._interp_table = np.load(self.init_props.power_grid_prop.base_power_props.path_datafile)
        self._sim = None
        self._sim_phase = None
        self._build()                                   # MA_DemandResponse.py:82

    # ---- build_environment (:84-119) ---------------------------------------------------------
    def _build(self) -> None:
        p = self.init_props
        hp, hv = p.cluster_prop.house_prop, p.cluster_prop.house_prop.hvac_prop
        n, mode = self.n, p.cluster_prop.agents_comm_prop.mode
        npz, caps = hp.noise_prop, hv.noise_prop.cooling_capacity_list
        st = {k: np.empty((1, n)) for k in ("target", "Ua", "Ca", "Cm", "Hm", "cap", "t_air", "t_mass")}
        self._house_props, self._hvac_props = [], []
        for i in range(n):                              # applyPropertyNoise, v0/utils.py:382-482
            b = hp.model_copy(deep=True)
            b.init_air_temp += abs(random.gauss(0, npz.std_start_temp))
            b.init_mass_temp += abs(random.gauss(0, npz.std_start_temp))
            b.target_temp += abs(random.gauss(0, npz.std_target_temp))
            b.Ua *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)   # multiplied in v0
            b.Cm *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)
            b.Ca *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)
            b.Hm *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)
            b.hvac_prop.cooling_capacity = random.choices(caps)[0]
            self._house_props.append(b)
            self._hvac_props.append(b.hvac_prop)
            st["target"][0, i], st["cap"][0, i] = b.target_temp, b.hvac_prop.cooling_capacity
            st["Ua"][0, i], st["Ca"][0, i], st["Cm"][0, i], st["Hm"][0, i] = b.Ua, b.Ca, b.Cm, b.Hm
            st["t_air"][0, i], st["t_mass"][0, i] = b.init_air_temp, b.init_mass_temp       # noised start (v0)
        self.start_datetime = p.start_datetime
        if p.start_datetime_mode == "random":           # v0/utils.py:509-518
            days, secs = random.randrange(364), random.randrange(60 * 60 * 24)
            self.start_datetime = p.start_datetime + _dt.timedelta(days=days, seconds=secs)
        self.date_time = self.datetime = self.start_datetime
        self.time_step = p.time_step
        # ClusterHouses.__init__ (:714-759): HVACs (lock-out noise draw), phase, first outdoor temperature, links
        ln = hv.noise_prop.lockout_noise
        # per-HVAC lock-out duration = lockout_duration + randint(-lockout_noise, lockout_noise) (:397-403)
        self._durations = [hv.lockout_duration + random.randint(-ln, ln) for _ in range(n)]
        if min(self._durations) < 1:
            raise ValueError("lockout_duration - lockout_noise must stay >= 1 s")
        tp = p.temp_prop
        phase = random.random() * 24 if tp.random_phase_offset else 0.0
        tp.phase = phase
        time_day = self.date_time.hour + self.date_time.minute / 60.0
        od = (tp.day_temp - tp.night_temp) / 2 * np.sin(2 * np.pi * (time_day + (-6 + phase)) / 24) \
            + (tp.day_temp + tp.night_temp) / 2
        od += random.gauss(0, tp.temp_std)
        self._default_max_consumption = hv.cooling_capacity / hv.cop
        self._max_power = 0.0
        for i in range(n):
            self._max_power += self._hvac_props[i].max_consumption      # noised capacities (:752-757)
        self._table = build_comm_table(p)               # random_fixed draws happen here (:803-808)
        self._width = comm_width(p)
        # PowerGrid.__init__ (:1037-1135)
        gp = p.power_grid_prop
        self._artificial_ratio = gp.artificial_ratio * gp.artificial_signal_ratio_range ** (random.random() * 2 - 1)
        self._perlin = None
        if gp.signal_properties.mode == "perlin":
            sp = gp.signal_properties
            self._perlin = _Perlin(sp.nb_octaves, sp.octaves_step, sp.period, random.random())
        if self._sim is None or self._sim_phase != phase:
            cfg = flatten_config(p, 1, self._precision, "hand_engineered", "external", "zero", 0, "auto", comm_table=True)
            self._sim = DrSim(cfg, self._device)
            self._sim_phase = phase
            if self._interp_table is not None:
                self._sim.set_interp_table(self._interp_table)
        st["on"] = np.zeros((1, n), dtype=np.uint8)      # HVACs start OFF, free to start (:401-403)
        st["lockout"] = np.zeros((1, n), dtype=np.uint8)
        st["sso"] = np.asarray([self._durations], dtype=np.int32)     # seconds_since_off = the HVAC's own duration
        st["lockout_duration"] = np.asarray([self._durations], dtype=np.int32)
        st.update(epoch=[to_epoch(self.date_time)], od_temp=[float(od)], signal=[0.0], base_power=[0.0],
                  artificial_ratio=[self._artificial_ratio], max_power=[self._max_power], solar=[0.0], power=[0.0],
                  t_since_interp=[gp.base_power_props.interp_update_period + 1])
        self._sim.set_state(st)
        # power_grid.step(start_datetime, time_step) (:117-119)
        ids = self._draw_interp_ids(will_fire=gp.base_power_props.mode == "interpolation")
        self._refresh(True, self._perlin_value(), ids)
        self._views()

    def reset(self):
        self._build()                                   # :131
        if self.init_props.cluster_prop.agents_comm_prop.mode == "random_sample":
            self._draw_sample_table()                   # make_cluster_obs_dict (:918-932)
            self._refresh(False)
        return self.get_obs()

    # ---- step (:144-180) ----------------------------------------------------------------------
    def step(self, action_dict: Dict[int, bool]):
        p = self.init_props
        self.date_time += p.time_step
        self.datetime = self.date_time
        actions = np.zeros((1, self.n), dtype=np.uint8)
        for i in range(self.n):                         # missing command -> False (:969-975)
            if i in action_dict and action_dict[i]:
                actions[0, i] = 1
        od_noise = np.asarray([random.gauss(0, p.temp_prop.temp_std)])         # compute_OD_temp (:1014)
        if p.cluster_prop.agents_comm_prop.mode == "random_sample":
            self._draw_sample_table()                   # one draw per step in v0 (:918-932)
            self._sim.set_comm_table(self._table)
        ids = self._draw_interp_ids(self._interp_will_fire())
        perlin = self._perlin_value()
        self._take(self._sim.step_host_snapshot(actions, od_noise, perlin, None if ids is None else ids))
        rew = self._reward
        obs = self.get_obs()
        rewards = {i: float(rew[i]) for i in range(self.n)}
        dones = {i: False for i in range(self.n)}       # :342-358
        return obs, rewards, dones, {"cluster_hvac_power": float(self._snap["power"][0])}

    # ---- legacy observation dicts (:868-940, :182-200) ------------------------------------------
    def get_obs(self) -> Dict[int, Dict[str, Any]]:
        s = self._snap
        mp = self.init_props.cluster_prop.message_prop
        n = self.n
        sso, on = s["sso"][0].tolist(), [bool(x) for x in s["on"][0].tolist()]
        lock = [bool(x) for x in s["lockout"][0].tolist()]
        ta, tm = s["t_air"][0].tolist(), s["t_mass"][0].tolist()
        solar, power, signal = float(s["solar"][0]), float(s["power"][0]), float(s["signal"][0])
        msgs = []
        for i in range(n):
            hv, b = self._hvac_props[i], self._house_props[i]
            m = {"current_temp_diff_to_target": ta[i] - b.target_temp,
                 "hvac_seconds_since_off": sso[i],
                 "hvac_curr_consumption": hv.max_consumption if on[i] else 0,
                 "hvac_max_consumption": hv.max_consumption,
                 "hvac_lockout_duration": self._durations[i]}
            if mp.thermal:
                m.update(house_Ua=b.Ua, house_Cm=b.Cm, house_Ca=b.Ca, house_Hm=b.Hm)
            if mp.hvac:
                m.update(hvac_COP=hv.cop, hvac_latent_cooling_fraction=hv.latent_cooling_fraction,
                         hvac_cooling_capacity=hv.cooling_capacity)
            msgs.append(m)
        table = self._table.tolist() if self._table.shape[1] else [[] for _ in range(n)]
        out: Dict[int, Dict[str, Any]] = {}
        for i in range(n):
            hv, b = self._hvac_props[i], self._house_props[i]
            out[i] = {
                "OD_temp": self.current_od_temp, "datetime": self.date_time,
                "house_temp": ta[i], "house_mass_temp": tm[i],
                "hvac_turned_on": on[i], "hvac_seconds_since_off": sso[i], "hvac_lockout": lock[i],
                "house_target_temp": b.target_temp, "house_deadband": b.deadband,
                "house_Ua": b.Ua, "house_Cm": b.Cm, "house_Ca": b.Ca, "house_Hm": b.Hm,
                "house_solar_gain": solar,
                "hvac_COP": hv.cop, "hvac_cooling_capacity": hv.cooling_capacity,
                "hvac_latent_cooling_fraction": hv.latent_cooling_fraction,
                "hvac_lockout_duration": self._durations[i],
                "message": [dict(msgs[j]) for j in table[i]],
                "reg_signal": signal, "cluster_hvac_power": power,
            }
        return out

    def __deepcopy__(self, memo):
        other = object.__new__(MADemandResponseEnv)
        memo[id(self)] = other
        for k, v in self.__dict__.items():
            if k in ("_sim", "cluster", "power_grid"):
                continue
            other.__dict__[k] = copy.deepcopy(v, memo)
        other._sim = self._sim.clone()
        other._views()
        return other
