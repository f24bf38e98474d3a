"""Drop-in replacement for ``app.core.environment.environment.Environment`` (environment.py:23).

Same constructor argument, ``reset()`` / ``step(action_dict)`` / ``get_obs()`` signatures, the
same 21-key per-agent observation dicts and reward dicts, the same attributes the reference's
callers read (``init_props``, ``date_time``, ``current_od_temp``, ``cluster.buildings[i]...``,
``cluster.max_power``, ``power_grid.current_signal``) and ``copy.deepcopy`` support
(training_manager.py:269).  Every step runs on the GPU through ``libdrsim.so``; the host only
draws the random numbers (from Python's global ``random``, in the reference's draw order, so
``random.seed(net_seed)`` reproduces the reference's trajectory) and materialises the dicts.
"""
from __future__ import annotations

import copy
import datetime as _dt
import random
from typing import Any, Dict, List, Optional

import numpy as np

from .core import DrSim, comm_width, flatten_config, from_epoch, nb_comm_of, to_epoch
from .properties import EnvironmentProperties, as_props

DAYS_IN_YEAR = 364.0  # environment.py:16
SECONDS_IN_DAY = 86400.0


# ------------------------------------------------------------------------------------------
# neighbour tables (agent_communication_builder.py:63-203)
# ------------------------------------------------------------------------------------------
def random_sample_ids(n_agents: int, agent_id: int, k: int) -> List[int]:
    """``get_random_sample`` (:191-203): ``random.sample`` over the ids without ``agent_id``."""
    ids = list(range(n_agents))
    ids.remove(agent_id)
    return random.sample(ids, k=k)


def build_comm_table(props: EnvironmentProperties) -> np.ndarray:
    """``get_comm_link_list`` as an ``[N, width]`` int32 array (``random_fixed`` draws from the
    global ``random``, one ``sample`` per agent in id order; ``random_sample`` -> zero width)."""
    cp = props.cluster_prop.agents_comm_prop
    n = props.cluster_prop.nb_agents
    c = nb_comm_of(props)
    rows: List[List[int]] = []
    if cp.mode == "neighbours":
        lo, hi = c // 2, (c + 1) // 2
        rows = [[(i - lo + j) % n for j in range(lo)] + [(i + 1 + j) % n for j in range(hi)] for i in range(n)]
    elif cp.mode == "closed_groups":
        for i in range(n):
            base = i - (i % (c + 1))
            if base + c <= n:
                ids = [base + j for j in range(cp.max_nb_agents_communication + 1)]
            else:
                ids = [n - c - 1 + j for j in range(c + 1)]
            ids.remove(i)
            rows.append(ids)
    elif cp.mode == "neighbours_2D":
        row, dist = cp.row_size, cp.max_communication_distance
        if n % row != 0:
            raise ValueError("Neighbours 2D row_size must be a divisor of nb_agents")
        max_y = n // row
        if dist >= (row + 1) // 2 or dist >= (max_y + 1) // 2:
            raise ValueError(
                "Neighbours 2D distance_comm ({}) must be strictly smaller than (row_size+1) / 2 ({}) and "
                "(max_y+1) / 2 ({})".format(dist, (row + 1) // 2, (max_y + 1) // 2))
        pattern = [(dx, dy) for dx in range(-dist, dist + 1) for dy in range(-dist, dist + 1)
                   if abs(dx) + abs(dy) <= dist and (dx != 0 or dy != 0)]
        for i in range(n):
            x, y = i % row, i // row
            rows.append([((y + dy) % max_y) * row + ((x + dx) % row) for dx, dy in pattern])
    elif cp.mode == "random_fixed":
        rows = [random_sample_ids(n, i, c) for i in range(n)]
    elif cp.mode == "random_sample":
        return np.zeros((n, 0), dtype=np.int32)
    else:
        raise AttributeError(f"unknown communication mode {cp.mode!r}")
    if len({len(r) for r in rows}) != 1:
        raise ValueError("ragged neighbour table (max_nb_agents_communication larger than the cluster)")
    return np.asarray(rows, dtype=np.int32).reshape(n, -1)


# ------------------------------------------------------------------------------------------
# perlin noise source (perlin.py:17-56); the third-party generator is used when importable
# ------------------------------------------------------------------------------------------
class _Perlin:
    def __init__(self, nb_octaves: int, octaves_step: int, period: float, seed: float):
        self.nb_octaves, self.period = nb_octaves, period
        try:
            from perlin_noise import PerlinNoise  # third-party, optional
        except Exception:  # noqa: BLE001
            PerlinNoise = None
        self.noise_list = None
        if PerlinNoise is not None:
            self.noise_list = [PerlinNoise(octaves=2 ** i * octaves_step, seed=seed) for i in range(nb_octaves)]

    @property
    def available(self) -> bool:
        return self.noise_list is not None

    def calculate_noise(self, x: float) -> float:
        noise = 0
        for j in range(self.nb_octaves - 1):
            noise += self.noise_list[j].noise(x / self.period) / (2 ** j)
        noise += self.noise_list[-1].noise(x / self.period) / (2 ** self.nb_octaves - 1)  # sic (quirk Q7)
        return 1 * noise


# ------------------------------------------------------------------------------------------
# attribute proxies the reference's callers expect
# ------------------------------------------------------------------------------------------
class _HvacView:
    def __init__(self, env: "Environment", i: int):
        self._e, self._i = env, i
        self.init_props = env._hvac_props[i]

    seconds_since_off = property(lambda s: int(s._e._snap["sso"][0, s._i]))
    lockout = property(lambda s: bool(s._e._snap["lockout"][0, s._i]))
    turned_on = property(lambda s: bool(s._e._snap["on"][0, s._i]))

    def get_power_consumption(self) -> float:
        return self.init_props.max_consumption if self.turned_on else 0.0


class _BuildingView:
    def __init__(self, env: "Environment", i: int):
        self._e, self._i = env, i
        self.init_props = env._house_props[i]
        self.hvac = _HvacView(env, i)
        self.max_consumption = env._default_max_consumption

    indoor_temp = property(lambda s: float(s._e._snap["t_air"][0, s._i]))
    current_mass_temp = property(lambda s: float(s._e._snap["t_mass"][0, s._i]))
    current_solar_gain = property(lambda s: float(s._e._snap["solar"][0]))

    def get_power_consumption(self) -> float:
        return self.hvac.get_power_consumption()


class _ClusterView:
    def __init__(self, env: "Environment"):
        self._e = env
        self.init_props = env.init_props.cluster_prop
        self.buildings = [_BuildingView(env, i) for i in range(env.n)]
        self.max_power = env._max_power

    current_power_consumption = property(lambda s: float(s._e._snap["power"][0]))

    @property
    def agent_communicators(self) -> Dict[int, List[int]]:
        t = self._e._table
        return {} if t is None or t.shape[1] == 0 else {i: [int(v) for v in t[i]] for i in range(t.shape[0])}


class _PowerGridView:
    def __init__(self, env: "Environment"):
        self._e = env
        self.init_props = env.init_props.power_grid_prop

    current_signal = property(lambda s: float(s._e._snap["signal"][0]))
    base_power = property(lambda s: float(s._e._snap["base_power"][0]))

    def get_obs(self):
        return {"reg_signal": self.current_signal}


class ObsDict(dict):
    """The per-agent observation dicts of ``Environment.get_obs`` (environment.py:110-130), plus the
    device-computed normalised vectors (``vectors``, what ``norm_state_dict`` returns).

    A real ``dict`` whose entries are MATERIALISED ON ACCESS: ``obs[i]`` builds house i's 21-key dict (and the
    message dicts of its neighbours) from the step's snapshot the first time it is asked for; anything that needs
    all of them (iteration, ``keys`` / ``items`` / ``values``, comparison, copy, pickling, ``dict(obs)``) fills the
    rest in id order first.  A policy that consumes ``norm_state_dict(obs, props)`` -- the vectors the GPU already
    gathered -- never pays for the 1000 x (21 + 10 x 5) Python objects of a large cluster.

    One caveat: C-level consumers that walk the dict storage directly (``json.dumps(obs)``) see only what has been
    materialised; hand them ``obs.materialize()`` (or build the environment with ``lazy_obs=False``)."""

    vectors: Optional[np.ndarray] = None

    def __init__(self, build=None, n: int = 0):
        super().__init__()
        self._build, self._n, self._full = build, int(n), build is None

    def __missing__(self, i):
        if self._build is None or isinstance(i, bool) or not isinstance(i, (int, np.integer)) or not 0 <= i < self._n:
            raise KeyError(i)
        d = self._build(int(i))
        dict.__setitem__(self, int(i), d)
        return d

    def _fill(self) -> None:
        if self._full:
            return
        have = {k: dict.__getitem__(self, k) for k in dict.keys(self)}
        dict.clear(self)
        for i in range(self._n):                       # id order, like the reference's loop (cluster.py:113-121)
            dict.__setitem__(self, i, have[i] if i in have else self._build(i))
        self._full = True

    def materialize(self) -> "ObsDict":
        """Build every entry now (id order); returns self."""
        self._fill()
        return self

    def __len__(self):
        return dict.__len__(self) if self._full else self._n

    def __contains__(self, i):
        if self._full:
            return dict.__contains__(self, i)
        return isinstance(i, (int, np.integer)) and not isinstance(i, bool) and 0 <= i < self._n

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def keys(self):
        self._fill()
        return dict.keys(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)

    def get(self, i, default=None):
        return self[i] if i in self else default

    def __eq__(self, other):
        self._fill()
        if isinstance(other, ObsDict):
            other._fill()
        return dict.__eq__(self, other)

    __hash__ = None

    def __repr__(self):
        self._fill()
        return dict.__repr__(self)

    def copy(self):
        self._fill()
        return dict(self)

    def __reduce__(self):
        self._fill()
        return (dict, (dict(dict.items(self)),))

    def __deepcopy__(self, memo):
        self._fill()
        return copy.deepcopy(dict(dict.items(self)), memo)


def norm_state_dict(obs_dicts: ObsDict, env_props: Any = None) -> List[np.ndarray]:
    """``app.utils.norm.norm_state_dict`` (norm.py:178-218) served from the device: the
    observation / neighbour-message gather already produced the normalised vectors."""
    v = getattr(obs_dicts, "vectors", None)
    if v is None:
        raise TypeError("norm_state_dict needs the ObsDict returned by marl_demandresponse_b200.Environment")
    return list(v)   # one row view per agent (a list, like the reference's)


class Environment:
    """GPU-backed stand-in for the reference ``Environment``."""

    def __init__(self, env_props: Any, device: int = 0, precision: str = "f64",
                 interp_table: Optional[np.ndarray] = None, lazy_obs: bool = True) -> None:
        self.init_props: EnvironmentProperties = as_props(env_props)  # deepcopy, environment.py:46
        self._lazy_obs = bool(lazy_obs)
        self.n = int(self.init_props.cluster_prop.nb_agents)
        self._device, self._precision = device, precision
        self._interp_table = interp_table
        if self.init_props.power_grid_prop.base_power_props.mode == "interpolation" and interp_table is None:
            self._interp_table = np.load(self.init_props.power_grid_prop.base_power_props.path_datafile)
        self._sim: Optional[DrSim] = None
        self._ids = list(range(self.n))
        self.reset()

    # ---- reset (environment.py:49-70, draw order of SURVEY appendix B) --------------------
    def _make_sim(self) -> DrSim:
        cfg = flatten_config(self.init_props, 1, self._precision, "hand_engineered", "external", "zero", 0, "auto",
                             comm_table=True)
        sim = DrSim(cfg, self._device)
        if self._interp_table is not None:
            sim.set_interp_table(self._interp_table)
        return sim

    def reset(self) -> ObsDict:
        p = self.init_props
        hp, hv = p.cluster_prop.house_prop, p.cluster_prop.house_prop.hvac_prop
        n, mode = self.n, p.cluster_prop.agents_comm_prop.mode
        if self._sim is None:
            self._sim = self._make_sim()
        # Cluster.reset (cluster.py:51-71): buildings, neighbour table, a first get_obs()
        self._default_max_consumption = hv.cooling_capacity / hv.cop
        self._max_power = 0.0
        for _ in range(n):
            self._max_power += self._default_max_consumption  # cached before the noise (quirk Q2)
        self._table = build_comm_table(p)
        self._width = comm_width(p)
        if mode == "random_sample":
            self._draw_sample_table()                      # the discarded get_obs of cluster.py:71
        # randomize_date (environment.py:176-194)
        self.date_time = p.start_datetime
        if p.start_datetime_mode == "random":
            days = random.randrange(int(DAYS_IN_YEAR))
            secs = random.randrange(int(SECONDS_IN_DAY))
            self.date_time = p.start_datetime + _dt.timedelta(days=days, seconds=secs)
        # Building.apply_noise / HVAC.apply_noise (building.py:224-267, hvac.py:66-70)
        self._house_props, self._hvac_props = [], []
        st = {k: np.empty((1, n)) for k in ("target", "Ua", "Ca", "Cm", "Hm", "cap")}
        for i in range(n):
            b = hp.model_copy(deep=True)
            npz = b.noise_prop
            b.init_air_temp += abs(random.gauss(0, npz.std_start_temp))
            b.init_mass_temp += abs(random.gauss(0, npz.std_start_temp))
            b.target_temp += abs(random.gauss(0, npz.std_target_temp))
            b.Ua = random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)   # '=' (quirk Q1)
            b.Cm *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)
            b.Ca *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)
            b.Hm *= random.triangular(npz.factor_thermo_low, npz.factor_thermo_high, 1)
            b.hvac_prop.cooling_capacity = random.choices(b.hvac_prop.noise_prop.cooling_capacity_list)[0]
            self._house_props.append(b)
            self._hvac_props.append(b.hvac_prop)
            st["target"][0, i], st["cap"][0, i] = b.target_temp, b.hvac_prop.cooling_capacity
            st["Ua"][0, i], st["Ca"][0, i], st["Cm"][0, i], st["Hm"][0, i] = b.Ua, b.Ca, b.Cm, b.Hm
        # Building.reset ran before the noise: un-noised initial temperatures, HVAC on (Q2, Q3)
        st["t_air"] = np.full((1, n), hp.init_air_temp)
        st["t_mass"] = np.full((1, n), hp.init_mass_temp)
        st["on"] = np.ones((1, n), dtype=np.uint8)
        st["lockout"] = np.zeros((1, n), dtype=np.uint8)
        st["sso"] = np.zeros((1, n), dtype=np.int32)
        # compute_od_temp (environment.py:132-159)
        tp = p.temp_prop
        amplitude = (tp.day_temp - tp.night_temp) / 2.0
        bias = (tp.day_temp + tp.night_temp) / 2.0
        time_day = self.date_time.hour + self.date_time.minute / 60.0
        od = amplitude * np.sin(2 * np.pi * (time_day + (-6.0 + tp.phase)) / 24.0) + bias
        od += random.gauss(0, tp.temp_std)
        # PowerGrid.__init__ (power_grid.py:41-66): the ratio draw compounds across resets because
        # the reference mutates the shared props object
        gp = p.power_grid_prop
        gp.artificial_ratio = gp.artificial_ratio * gp.artificial_signal_ratio_range ** (random.random() * 2 - 1)
        self._perlin = None
        if gp.signal_properties.mode == "perlin":          # signal_calculator.py:24-31
            sp = gp.signal_properties
            self._perlin = _Perlin(sp.nb_octaves, sp.octaves_step, sp.period, random.random())
        st.update(epoch=[to_epoch(self.date_time)], od_temp=[float(od)], signal=[0.0], base_power=[0.0],
                  artificial_ratio=[gp.artificial_ratio], max_power=[self._max_power], solar=[0.0],
                  power=[n * self._default_max_consumption],   # cached pre-noise (cluster.py:61-64)
                  t_since_interp=[gp.base_power_props.interp_update_period + 1])
        self._sim.set_state(st)
        # power_grid.step at reset (environment.py:66-68), then get_obs
        ids = self._draw_interp_ids(will_fire=gp.base_power_props.mode == "interpolation")
        perlin = self._perlin_value()
        if mode == "random_sample":
            self._draw_sample_table()
        self._refresh(True, perlin, ids)
        self._views()
        return self.get_obs()

    # ---- helpers --------------------------------------------------------------------------
    def _views(self) -> None:
        self.cluster = _ClusterView(self)
        self.power_grid = _PowerGridView(self)

    def _draw_sample_table(self) -> None:
        self._table = np.asarray([random_sample_ids(self.n, i, self._width) for i in range(self.n)],
                                 dtype=np.int32).reshape(self.n, -1)

    def _draw_interp_ids(self, will_fire: bool):
        bp = self.init_props.power_grid_prop.base_power_props
        if will_fire and self.n > bp.interp_nb_agents:
            return np.asarray(random.choices(list(range(self.n)), k=bp.interp_nb_agents), dtype=np.int32)[None]
        return None

    def _perlin_value(self):
        if self._perlin is None or not self._perlin.available:
            return None
        import time

        x = time.mktime(self.date_time.timetuple()) % 86400      # signal_calculator.py:113
        return np.asarray([self._perlin.calculate_noise(x)], dtype=np.float64)

    def _interp_will_fire(self) -> bool:
        bp = self.init_props.power_grid_prop.base_power_props
        if bp.mode != "interpolation":
            return False
        return self._tsi + int(self.init_props.time_step.seconds) >= bp.interp_update_period

    def _upload(self, x, dtype):
        import torch

        return None if x is None else torch.as_tensor(np.ascontiguousarray(x, dtype=dtype), device=f"cuda:{self._device}")

    def _refresh(self, recompute_signal: bool, perlin=None, ids=None) -> None:
        if self._width > 0:
            self._sim.set_comm_table(self._table)
        self._sim.refresh(recompute_signal, None, self._upload(perlin, np.float64), self._upload(ids, np.int32))
        self._take(self._sim.snapshot())

    def _take(self, snap: Dict[str, np.ndarray]) -> None:
        """Keep the step's snapshot (``drsim_snapshot``: one kernel writing pinned host memory + one copy of the
        observation rows, ONE synchronisation): the pinned buffers are reused by the next step, callers hold on
        to observation dicts of earlier steps, so everything is copied out here -- a few KB."""
        env = snap["env"][0].copy()
        if getattr(self, "_targets_of", None) is not self._house_props:     # static per reset
            self._targets = np.asarray([[b.target_temp for b in self._house_props]], dtype=np.float64)
            self._caps = np.asarray([[h.cooling_capacity for h in self._hvac_props]], dtype=np.float64)
            self._targets_of = self._house_props
        self._snap = {"target": self._targets, "cap": self._caps, "epoch": np.asarray([int(env[5])], dtype=np.int64),
                      "t_air": snap["t_air"].copy(), "t_mass": snap["t_mass"].copy(), "sso": snap["sso"].copy(),
                      "on": snap["on"].copy(), "lockout": snap["lockout"].copy(),
                      "od_temp": env[0:1], "signal": env[1:2], "power": env[2:3], "solar": env[3:4], "base_power": env[4:5]}
        self._reward = snap["reward"][0].copy()
        self._tsi = int(env[6])
        self._vectors = None if snap["obs"] is None else snap["obs"][0].astype(np.float64)
        self.current_od_temp = float(env[0])

    # ---- step (environment.py:72-108) -----------------------------------------------------
    def step(self, action_dict: Dict[int, bool]):
        p = self.init_props
        mode = p.cluster_prop.agents_comm_prop.mode
        self.date_time += p.time_step
        n = self.n
        if len(action_dict) == n and list(action_dict) == self._ids:      # the usual case: one command per house, id order
            actions = np.fromiter(action_dict.values(), dtype=np.bool_, count=n).view(np.uint8).reshape(1, n)
        else:
            actions = np.zeros((1, n), dtype=np.uint8)
            for i in range(n):                                  # missing action -> False (cluster.py:83-86)
                if i in action_dict and action_dict[i]:
                    actions[0, i] = 1
        if mode == "random_sample":
            self._draw_sample_table()                           # discarded get_obs of cluster.py:89 (quirk Q4)
        od_noise = np.asarray([random.gauss(0, p.temp_prop.temp_std)])
        ids = self._draw_interp_ids(self._interp_will_fire())
        perlin = self._perlin_value()
        if mode == "random_sample":
            self._draw_sample_table()
            self._sim.set_comm_table(self._table)
        self._take(self._sim.step_host_snapshot(actions, od_noise, perlin, None if ids is None else ids))
        return self.get_obs(), dict(enumerate(self._reward.tolist()))

    # ---- observations (environment.py:110-130, cluster.py:91-121, building.py:79-139) ------
    def get_obs(self) -> ObsDict:
        s = self._snap
        mp = self.init_props.cluster_prop.message_prop
        n = self.n
        hvac_props, house_props = self._hvac_props, self._house_props
        table = self._table
        date_time, od_temp = self.date_time, self.current_od_temp
        solar, power, signal = float(s["solar"][0]), float(s["power"][0]), float(s["signal"][0])
        lists: Dict[str, list] = {}
        msgs: Dict[int, dict] = {}

        def col(k):
            # plain Python lists (per-element numpy indexing would dominate the dict building), made on first use
            v = lists.get(k)
            if v is None:
                v = s[k][0].tolist()
                if k in ("on", "lockout"):
                    v = [bool(x) for x in v]
                lists[k] = v
            return v

        def message(j):                                         # Building.message (building.py:102-139)
            m = msgs.get(j)
            if m is None:
                hp_j, b = hvac_props[j], house_props[j]
                pmax = hp_j.max_consumption
                m = {
                    "seconds_since_off": col("sso")[j],
                    "curr_consumption": pmax if col("on")[j] else 0.0,
                    "max_consumption": pmax,
                    "lockout_duration": hp_j.lockout_duration,
                    "current_temp_diff_to_target": col("t_air")[j] - b.target_temp,
                }
                if mp.hvac:
                    m.update(cop=hp_j.cop, latent_cooling_fraction=hp_j.latent_cooling_fraction,
                             cooling_capacity=hp_j.cooling_capacity)
                if mp.thermal:
                    m.update(Ca=b.Ca, Ua=b.Ua, Cm=b.Cm, Hm=b.Hm)
                msgs[j] = m
            return m

        def build(i):
            hp_i, b = hvac_props[i], house_props[i]
            return {
                "turned_on": col("on")[i],
                "seconds_since_off": col("sso")[i],
                "lockout": col("lockout")[i],
                "cop": hp_i.cop,
                "cooling_capacity": hp_i.cooling_capacity,
                "latent_cooling_fraction": hp_i.latent_cooling_fraction,
                "lockout_duration": hp_i.lockout_duration,
                "target_temp": b.target_temp,
                "deadband": b.deadband,
                "Ua": b.Ua, "Ca": b.Ca, "Cm": b.Cm, "Hm": b.Hm,
                "indoor_temp": col("t_air")[i],
                "mass_temp": col("t_mass")[i],
                "solar_gain": solar,
                "cluster_hvac_power": power,
                "message": [dict(message(int(j))) for j in table[i]] if table.shape[1] else [],
                "OD_temp": od_temp,
                "datetime": date_time,
                "reg_signal": signal,
            }

        out = ObsDict(build, n)
        out.vectors = self._vectors
        if not getattr(self, "_lazy_obs", True):
            out.materialize()
        return out

    # ---- deepcopy (training_manager.py:269) -----------------------------------------------
    def __deepcopy__(self, memo):
        other = object.__new__(Environment)
        memo[id(self)] = other
        for k, v in self.__dict__.items():
            if k in ("_sim", "cluster", "power_grid"):
                continue
            other.__dict__[k] = copy.deepcopy(v, memo)
        other._sim = self._sim.clone()
        other._views()
        return other
