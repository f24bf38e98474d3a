"""Generate the golden trajectories under ``tests/golden/*.npz`` from the REAL reference.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Each case resets the reference ``Environment`` from a Python ``random`` seed, records the full
post-reset state, then steps it ``T`` times while logging every injected quantity (actions,
the outdoor-temperature gauss draw, the perlin value, the interpolator's sampled ids, the
per-call ``random_sample`` neighbour tables) and every output (house state, env scalars,
rewards, ``norm_state_dict`` vectors).  The CUDA path, the NumPy oracle and the scalar port
are all checked against these files; nothing here is read from ``/root/reference`` at test
time.  The interpolation table is the seeded synthetic one (``oracle.config.synthetic_table``)
because the real ``mergedGridSearchResultFinal.npy`` is missing from the reference checkout.
"""
from __future__ import annotations

import copy
import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refenv  # noqa: E402
from oracle.config import normalize_env_prop, synthetic_table  # noqa: E402
from oracle.np_oracle import deadband_bangbang, to_epoch  # noqa: E402

MARL_JSON = os.path.join(refenv.REFERENCE_ROOT, "server/app/core/config/MARLconfig.json")
MC_DIR = os.path.join(refenv.REFERENCE_ROOT, "server/v0/monteCarlo")
TABLE_SEED = 2024


def base_cfg(n, **over):
    cfg = json.load(open(MARL_JSON))["env_prop"]
    cfg["cluster_prop"]["nb_agents"] = n
    for path, v in over.items():
        d = cfg
        keys = path.split("/")
        for k in keys[:-1]:
            d = d[k]
        d[keys[-1]] = v
    return cfg


CASES = [
    # name, cfg, T, policy, seed
    dict(name="c1_default_n10_bangbang", cfg=base_cfg(10), T=150, policy="bangbang", seed=4),
    dict(name="random_n24_sinus_commonL2", T=120, policy="random", seed=11, cfg=base_cfg(
        24, **{"power_grid_prop/signal_properties/mode": "sinusoidals",
               "reward_prop/penalty_props/mode": "common_L2",
               "cluster_prop/house_prop/deadband": 1.0})),
    dict(name="steps_n22_commonmax_closedgroups_allflags", T=100, policy="random", seed=12, cfg=base_cfg(
        22, **{"power_grid_prop/signal_properties/mode": "regular_steps",
               "reward_prop/penalty_props/mode": "common_max_error",
               "cluster_prop/agents_comm_prop/mode": "closed_groups",
               "cluster_prop/agents_comm_prop/max_nb_agents_communication": 4,
               "state_prop/thermal": True, "state_prop/hvac": True, "state_prop/solar_gain": True,
               "state_prop/hour": True, "state_prop/day": True,
               "cluster_prop/message_prop/thermal": True, "cluster_prop/message_prop/hvac": True,
               "start_datetime": "2021-03-01T06:40:00", "start_datetime_mode": "fixed"})),
    dict(name="flat_n25_mixture_2d_nosolar", T=100, policy="bangbang", seed=13, cfg=base_cfg(
        25, **{"power_grid_prop/signal_properties/mode": "flat",
               "reward_prop/penalty_props/mode": "mixture",
               "reward_prop/penalty_props/alpha_common_max": 0.5,
               "reward_prop/alpha_sig": 0.7, "reward_prop/alpha_temp": 1.3,
               "cluster_prop/agents_comm_prop/mode": "neighbours_2D",
               "cluster_prop/agents_comm_prop/max_communication_distance": 2,
               "cluster_prop/house_prop/solar_gain": False,
               "power_grid_prop/artificial_ratio": 0.9,
               "power_grid_prop/artificial_signal_ratio_range": 2})),
    dict(name="interp_n12_all", T=170, policy="bangbang", seed=14, table=True, cfg=base_cfg(
        12, **{"power_grid_prop/base_power_props/mode": "interpolation",
               "power_grid_prop/signal_properties/mode": "flat"})),
    dict(name="interp_n130_sampled", T=160, policy="random", seed=15, table=True, obs_stride=10, cfg=base_cfg(
        130, **{"power_grid_prop/base_power_props/mode": "interpolation",
                "start_datetime": "2021-06-15T11:58:00", "start_datetime_mode": "fixed"})),
    dict(name="randomfixed_n16", T=60, policy="random", seed=16, cfg=base_cfg(
        16, **{"cluster_prop/agents_comm_prop/mode": "random_fixed",
               "cluster_prop/agents_comm_prop/max_nb_agents_communication": 6})),
    dict(name="randomsample_n9", T=40, policy="random", seed=17, cfg=base_cfg(
        9, **{"cluster_prop/agents_comm_prop/mode": "random_sample",
              "cluster_prop/agents_comm_prop/max_nb_agents_communication": 3})),
    dict(name="newyear_rollover_n8", T=90, policy="bangbang", seed=18, cfg=base_cfg(
        8, **{"start_datetime": "2023-12-31T23:57:00", "start_datetime_mode": "fixed",
              "cluster_prop/house_prop/hvac_prop/lockout_duration": 12})),
    dict(name="solar_edge_n8", T=90, policy="random", seed=19, cfg=base_cfg(
        8, **{"start_datetime": "2024-02-29T07:27:00", "start_datetime_mode": "fixed",
              "time_step": 7.0, "temp_prop/phase": 1.5, "temp_prop/day_temp": 31.0})),
    dict(name="greedy_n40", T=100, policy="greedy", seed=20, cfg=base_cfg(
        40, **{"power_grid_prop/signal_properties/mode": "sinusoidals"})),
]


def greedy_actions(ns, obs):
    """Drive the reference's own GreedyMyopic (greedy_myopic_controller.py:67-104)."""
    gm = refenv.load_controller_module("greedy_myopic_controller")

    ctl = gm.GreedyMyopic({"id": 0}, None)
    ctl.get_action(obs)
    df = gm.GreedyMyopic.actions_df
    return np.array([bool(df.loc[i]["HVAC_status"]) for i in range(len(obs))])


def run_case(ns, case, table_path):
    cfg = copy.deepcopy(case["cfg"])
    if case.get("table"):
        bp = cfg["power_grid_prop"]["base_power_props"]
        bp["path_datafile"] = table_path
        bp["path_parameter_dict"] = os.path.join(MC_DIR, "interp_parameters_dict.json")
        bp["path_dict_keys"] = os.path.join(MC_DIR, "interp_dict_keys.csv")
    N = cfg["cluster_prop"]["nb_agents"]
    T = case["T"]
    random.seed(case["seed"])
    rng = np.random.default_rng(case["seed"])
    out = {}
    with refenv.Recorder(ns) as rec:
        env = ns.Environment(ns.EnvironmentProperties(**cfg))
        # Environment.__init__ already reset once (environment.py:46-47); reset again so the
        # recorded draw stream is the one of a plain ``reset()`` after ``random.seed``.
        random.seed(case["seed"])
        n_od, n_pe, n_in, n_cs = len(rec.od_noise), len(rec.perlin), len(rec.interp_ids), len(rec.comm_samples)
        obs = env.reset()
        st0 = refenv.extract_state(env)
        for k, v in st0.items():
            out["state0_" + k] = v
        out["reset_od_noise"] = np.array(rec.od_noise[n_od:])
        out["reset_perlin"] = np.array(rec.perlin[n_pe:])
        k_interp = cfg["power_grid_prop"]["base_power_props"]["interp_nb_agents"]
        out["reset_interp_ids"] = np.array(rec.interp_ids[n_in:], dtype=np.int32).reshape(-1, k_interp)
        mode = cfg["cluster_prop"]["agents_comm_prop"]["mode"]
        per_step_comm = mode == "random_sample"
        if per_step_comm:
            comm = [np.array(rec.comm_samples[-N:], dtype=np.int32)]
        else:
            out["comm_table"] = np.array([env.cluster.agent_communicators[i] for i in range(N)], dtype=np.int32)
        vec = [np.array(ns.norm_state_dict(obs, env.init_props))]
        keys_f = ("t_air", "t_mass")
        keys_e = ("power", "signal", "od_temp", "solar", "base_power")
        log = {k: [] for k in keys_f + keys_e + ("on", "lockout", "sso", "rewards", "actions",
                                                  "od_noise", "perlin", "interp_ids", "epoch", "t_since_interp")}
        k_interp = cfg["power_grid_prop"]["base_power_props"]["interp_nb_agents"]
        for t in range(T):
            if case["policy"] == "random":
                a = rng.random(N) < 0.5
            elif case["policy"] == "bangbang":
                a = deadband_bangbang(
                    np.array([obs[i]["indoor_temp"] for i in range(N)]),
                    np.array([obs[i]["target_temp"] for i in range(N)]),
                    np.array([obs[i]["deadband"] for i in range(N)]),
                    np.array([obs[i]["turned_on"] for i in range(N)]))
            elif case["policy"] == "greedy":
                a = greedy_actions(ns, obs)
            n_od, n_pe, n_in = len(rec.od_noise), len(rec.perlin), len(rec.interp_ids)
            obs, rew = env.step({i: bool(a[i]) for i in range(N)})
            st = refenv.extract_state(env)
            log["actions"].append(a.astype(np.uint8))
            log["od_noise"].append(rec.od_noise[n_od])
            log["perlin"].append(rec.perlin[n_pe] if len(rec.perlin) > n_pe else np.nan)
            ids = rec.interp_ids[n_in] if len(rec.interp_ids) > n_in else [-1] * k_interp
            log["interp_ids"].append(np.array(ids, dtype=np.int32))
            for k in keys_f:
                log[k].append(st[k][0])
            for k in keys_e:
                log[k].append(st[k][0])
            log["on"].append(st["on"][0].astype(np.uint8))
            log["lockout"].append(st["lockout"][0].astype(np.uint8))
            log["sso"].append(st["sso"][0].astype(np.int32))
            log["epoch"].append(st["epoch"][0])
            log["t_since_interp"].append(st.get("t_since_interp", np.array([-1]))[0])
            log["rewards"].append(np.array([rew[i] for i in range(N)]))
            vec.append(np.array(ns.norm_state_dict(obs, env.init_props)))
            if per_step_comm:
                comm.append(np.array(rec.comm_samples[-N:], dtype=np.int32))
        for k, v in log.items():
            out[k] = np.array(v)
        stride = case.get("obs_stride", 1)
        out["obs"] = np.array(vec)[::stride]
        if per_step_comm:
            out["comm_per_step"] = np.array(comm)
        # raw dict observation of house 0 at the last step, for the dict-API test
        o = dict(obs[0])
        o["datetime"] = o["datetime"].isoformat()
        o = json.loads(json.dumps(o, default=float))
        out["last_obs_house0_json"] = np.frombuffer(json.dumps(o).encode(), dtype=np.uint8)
    cfg_store = copy.deepcopy(case["cfg"])
    meta = dict(name=case["name"], seed=case["seed"], T=T, policy=case["policy"],
                obs_stride=case.get("obs_stride", 1),
                table_seed=TABLE_SEED if case.get("table") else None, env_prop=cfg_store)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def main():
    ns = refenv.load()
    tmp = tempfile.mkdtemp()
    table_path = os.path.join(tmp, "table.npy")
    np.save(table_path, synthetic_table(TABLE_SEED))
    for case in CASES:
        out = run_case(ns, case, table_path)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']:48s} {os.path.getsize(path)/1024:8.1f} KiB  obs{out['obs'].shape}")


if __name__ == "__main__":
    main()
