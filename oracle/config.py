"""Default ``env_prop`` tree + normalisation (TEST INFRASTRUCTURE, see ``oracle/__init__``).

The defaults restate the pydantic defaults of the reference so the oracle is usable on the
GPU box where ``/root/reference`` is absent:

* ``server/app/core/environment/environment_properties.py:51-54,70-89,115-130,144-205,
  221-258,270-311,325-336,355-366``
* ``server/app/core/environment/cluster/cluster_properties.py:9-39``
* ``server/app/core/environment/power_grid/power_grid_properties.py:21-27,46-53,67-70``
"""
from __future__ import annotations

import copy
import datetime as _dt

DEFAULT_ENV_PROP = {
    "start_datetime": "2021-01-01T00:00:00",
    "start_datetime_mode": "random",
    "time_step": 4.0,
    "temp_prop": {
        "day_temp": 26.0,
        "night_temp": 20.0,
        "temp_std": 1.0,
        "random_phase_offset": False,
        "phase": 0.0,
    },
    "state_prop": {"hour": False, "day": False, "solar_gain": False, "thermal": False, "hvac": False},
    "reward_prop": {
        "alpha_temp": 1.0,
        "alpha_sig": 1.0,
        "norm_reg_sig": 7500,
        "penalty_props": {
            "mode": "individual_L2",
            "alpha_ind_l2": 1.0,
            "alpha_common_l2": 1.0,
            "alpha_common_max": 0.0,
        },
        "sig_penalty_mode": "common_L2",
    },
    "cluster_prop": {
        "nb_agents": 1000,
        "nb_agents_comm": 10,
        "agents_comm_prop": {
            "mode": "neighbours",
            "row_size": 5,
            "max_communication_distance": 2,
            "max_nb_agents_communication": 10,
        },
        "message_prop": {"thermal": False, "hvac": False},
        "house_prop": {
            "Ua": 2.18e02,
            "Ca": 9.08e05,
            "Hm": 2.84e03,
            "Cm": 3.45e06,
            "target_temp": 20.0,
            "deadband": 0.0,
            "init_air_temp": 20.0,
            "init_mass_temp": 20.0,
            "solar_gain": True,
            "window_area": 7.175,
            "shading_coeff": 0.67,
            "noise_prop": {
                "std_start_temp": 3.0,
                "std_target_temp": 1.0,
                "factor_thermo_low": 0.9,
                "factor_thermo_high": 1.1,
            },
            "hvac_prop": {
                "cop": 2.5,
                "cooling_capacity": 15000.0,
                "latent_cooling_fraction": 0.35,
                "lockout_duration": 40,
                "noise_prop": {
                    "std_latent_cooling_fraction": 0.05,
                    "factor_COP_low": 0.95,
                    "factor_COP_high": 1.05,
                    "factor_cooling_capacity_low": 0.9,
                    "factor_cooling_capacity_high": 1.1,
                    "lockout_noise": 0,
                    "cooling_capacity_list": [12500, 15000, 17500],
                },
            },
        },
    },
    "power_grid_prop": {
        "artificial_signal_ratio_range": 1,
        "artificial_ratio": 1.0,
        "base_power_props": {
            "mode": "constant",
            "avg_power_per_hvac": 4200,
            "init_signal_per_hvac": 910,
            "path_datafile": "./monteCarlo/mergedGridSearchResultFinal.npy",
            "path_parameter_dict": "./monteCarlo/interp_parameters_dict.json",
            "path_dict_keys": "./monteCarlo/interp_dict_keys.csv",
            "interp_update_period": 300,
            "interp_nb_agents": 100,
        },
        "signal_properties": {
            "mode": "perlin",
            "amplitude_ratios": [0.1, 0.3],
            "amplitude_per_hvac": 6000,
            "nb_octaves": 5,
            "octaves_step": 5,
            "period": 300,
            "periods": [400, 1200],
        },
    },
}

# server/v0/monteCarlo/interp_parameters_dict.json:1 and interp_dict_keys.csv:1
INTERP_KEYS = [
    "Ua_ratio", "Cm_ratio", "Ca_ratio", "Hm_ratio",
    "air_temp", "mass_temp", "OD_temp", "HVAC_power", "hour", "date",
]
INTERP_GRIDS = {
    "Ua_ratio": [0.9, 1, 1.1],
    "Cm_ratio": [0.9, 1, 1.1],
    "Ca_ratio": [0.9, 1, 1.1],
    "Hm_ratio": [0.9, 1, 1.1],
    "air_temp": [-4, -2, -1, -0.3, 0, 0.3, 1, 2, 4],
    "mass_temp": [-4, -2, 0, 2, 4],
    "OD_temp": [1, 3, 5, 7, 9, 11, 13, 15],
    "HVAC_power": [10000, 15000],
    "hour": [0.0, 10800.0, 21600.0, 25200.0, 27000.0, 39600.0, 46800.0, 57600.0,
             61200.0, 63000.0, 75600.0, 86399.0],
    "date": [0, 79, 171, 263, 354, 364],
}
INTERP_SHAPE = tuple(len(INTERP_GRIDS[k]) for k in INTERP_KEYS)  # (3,3,3,3,9,5,8,2,12,6)


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def normalize_env_prop(env_prop: dict | None) -> dict:
    """Deep-merge ``env_prop`` over the reference defaults; parse datetime / timestep."""
    out = copy.deepcopy(DEFAULT_ENV_PROP)
    if env_prop:
        _merge(out, env_prop)
    sd = out["start_datetime"]
    if isinstance(sd, str):
        sd = _dt.datetime.fromisoformat(sd)
    out["start_datetime"] = sd
    ts = out["time_step"]
    if isinstance(ts, _dt.timedelta):
        ts = ts.seconds  # the reference only ever reads ``time_step.seconds`` (hvac.py:46)
    out["time_step"] = int(ts)
    return out


def synthetic_table(seed: int = 2024):
    """Seeded stand-in for the missing ``mergedGridSearchResultFinal.npy``."""
    import numpy as np

    n = 1
    for s in INTERP_SHAPE:
        n *= s
    return np.random.default_rng(seed).uniform(0.0, 6000.0, n)
