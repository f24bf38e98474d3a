"""Monte-Carlo interpolation-table generator on the GPU.

Restates ``server/v0/monteCarlo/monteCarlo.py:152-230``: for each of the 4,199,040 grid points of
``interp_parameters_dict.json`` one noise-free single-house cluster is simulated for 75 steps under
the ``BangBangController`` (``v0/agents/bangbang_controllers.py:39-59``) -- thermal parameters =
defaults x the four ratios, initial air / mass temperature = target + offset, constant outdoor
temperature = target + ``OD_temp`` (``temp_mode = "constant"``, ``temp_std = 0``), fixed start
datetime (``date`` days after 2021-01-01 at ``hour`` seconds), lock-out 1 s, HVAC initially off with
``seconds_since_off = lockout_duration`` (``v0/env/MA_DemandResponse.py:397-403``) -- and the table
entry is the mean of the last 10 running averages of the HVAC power (monteCarlo.py:216-224).

The reference needs hours of Python for this (and its result file is missing from the checkout); here
every grid point is one replica of ``BatchedEnv`` with the on-device bang-bang policy.
"""
from __future__ import annotations

import datetime as _dt
import itertools
from typing import Dict, Optional, Sequence

import numpy as np

from .batched import BatchedEnv
from .core import to_epoch

# monteCarlo.py:73-124 (hours already in seconds, dates as days after 2021-01-01)
GRID: Dict[str, Sequence[float]] = {
    "Ua_ratio": [0.9, 1, 1.1], "Cm_ratio": [0.9, 1, 1.1], "Ca_ratio": [0.9, 1, 1.1], "Hm_ratio": [0.9, 1, 1.1],
    "air_temp": [-4, -2, -1, -0.3, 0, 0.3, 1, 2, 4], "mass_temp": [-4, -2, 0, 2, 4],
    "OD_temp": [1, 3, 5, 7, 9, 11, 13, 15], "HVAC_power": [10000, 15000],
    "hour": [0.0, 10800.0, 21600.0, 25200.0, 27000.0, 39600.0, 46800.0, 57600.0, 61200.0, 63000.0, 75600.0, 86399.0],
    "date": [0, 79, 171, 263, 354, 364],
}
KEYS = list(GRID.keys())
NB_TIME_STEPS_BY_SIM = 75   # monteCarlo.py:18
NB_TIME_STEPS_AVG = 10      # monteCarlo.py:19


def grid_points(index: np.ndarray) -> Dict[str, np.ndarray]:
    """Flat table index (row-major, last key fastest: ``unit_tests_interp.py:15-44``) -> parameters."""
    shape = tuple(len(GRID[k]) for k in KEYS)
    sub = np.unravel_index(np.asarray(index, dtype=np.int64), shape)
    return {k: np.asarray(GRID[k], dtype=np.float64)[i] for k, i in zip(KEYS, sub)}


def env_prop_for_table(house: Optional[dict] = None) -> dict:
    hp = {"Ua": 218.0, "Ca": 9.08e5, "Cm": 3.45e6, "Hm": 2.84e3, "target_temp": 20.0, "deadband": 0.0,
          "solar_gain": True, "window_area": 7.175, "shading_coeff": 0.67,
          "hvac_prop": {"cop": 2.5, "cooling_capacity": 15000.0, "latent_cooling_fraction": 0.35, "lockout_duration": 1}}
    if house:
        hp.update(house)
    t = hp["target_temp"]
    return {
        "start_datetime": "2021-01-01T00:00:00", "start_datetime_mode": "fixed", "time_step": 4.0,
        # outdoor temperature = target + OD_temp, injected per replica as the "noise" term
        "temp_prop": {"day_temp": t, "night_temp": t, "temp_std": 0.0, "phase": 0.0},
        "cluster_prop": {"nb_agents": 1, "house_prop": hp,
                         "agents_comm_prop": {"mode": "neighbours", "max_nb_agents_communication": 0}},
        "power_grid_prop": {"base_power_props": {"mode": "constant"}, "signal_properties": {"mode": "flat"}},
    }


def initial_state(pts: Dict[str, np.ndarray], prop: dict) -> Dict[str, np.ndarray]:
    hp = prop["cluster_prop"]["house_prop"]
    R = len(pts["date"])
    t = hp["target_temp"]
    col = lambda v: np.asarray(v, dtype=np.float64).reshape(R, 1)
    d0 = to_epoch(_dt.datetime(2021, 1, 1))
    st = {
        "target": np.full((R, 1), t), "t_air": col(t + pts["air_temp"]), "t_mass": col(t + pts["mass_temp"]),
        "Ua": col(hp["Ua"] * pts["Ua_ratio"]), "Cm": col(hp["Cm"] * pts["Cm_ratio"]),
        "Ca": col(hp["Ca"] * pts["Ca_ratio"]), "Hm": col(hp["Hm"] * pts["Hm_ratio"]), "cap": col(pts["HVAC_power"]),
        "on": np.zeros((R, 1), dtype=np.uint8), "lockout": np.zeros((R, 1), dtype=np.uint8),
        "sso": np.full((R, 1), hp["hvac_prop"]["lockout_duration"], dtype=np.int32),
        # int(hour // 3600) h, int(hour % 3600 // 60) min, int(hour % 60) s  (monteCarlo.py:166-168)
        "epoch": (d0 + pts["date"].astype(np.int64) * 86400 + (pts["hour"] // 3600).astype(np.int64) * 3600
                  + (pts["hour"] % 3600 // 60).astype(np.int64) * 60 + (pts["hour"] % 60).astype(np.int64)),
        "od_temp": t + pts["OD_temp"], "signal": np.zeros(R), "base_power": np.zeros(R), "power": np.zeros(R),
        "solar": np.zeros(R), "artificial_ratio": np.ones(R), "max_power": np.full(R, 1e12),
    }
    return st


def generate_table(indices: Optional[np.ndarray] = None, device: int = 0, precision: str = "f32",
                   chunk: int = 1 << 19, house: Optional[dict] = None) -> np.ndarray:
    """Average bang-bang HVAC power for the grid points ``indices`` (default: the whole table, in the
    reference's flat order, ready for ``PowerInterpolator`` / ``DrSim.set_interp_table``)."""
    import torch

    n_total = int(np.prod([len(GRID[k]) for k in KEYS]))
    if indices is None:
        indices = np.arange(n_total, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    prop = env_prop_for_table(house)
    out = np.empty(indices.shape[0], dtype=np.float64)
    env = None
    for lo in range(0, indices.shape[0], chunk):
        idx = indices[lo:lo + chunk]
        pts = grid_points(idx)
        R = idx.shape[0]
        if env is None or env.n_replicas != R:
            env = BatchedEnv(prop, R, device=device, precision=precision, obs_layout="none", policy="bangbang",
                             noise="zero")
        env.set_state(initial_state(pts, prop))
        od = torch.as_tensor(pts["OD_temp"], dtype=torch.float64, device=f"cuda:{device}")  # constant outdoor temp
        total = torch.zeros(R, dtype=torch.float64, device=od.device)
        avg = torch.zeros_like(total)
        for i in range(NB_TIME_STEPS_BY_SIM):
            env.step(None, od_noise=od)
            total += env.power
            if i >= NB_TIME_STEPS_BY_SIM - NB_TIME_STEPS_AVG:
                avg += total / ((i + 1) * NB_TIME_STEPS_AVG)
        out[lo:lo + R] = avg.cpu().numpy()
    return out
