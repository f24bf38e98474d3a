"""Generate ``tests/golden/v0_*.npz`` from the REAL legacy environment of the reference
(``server/v0/env/MA_DemandResponse.py``), for the v0 adapter.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_golden_v0.py

The legacy module is imported with stubs for what this image lacks (``ray``, ``matplotlib``,
``perlin_noise``, ``wandb``).  Each case seeds Python's ``random``, constructs the env, calls
``reset()`` and steps it ``T`` times with actions that are a pure function of (step, agent), logging
the observation scalars of every agent, rewards and ``info``; the adapter test replays the same
seed / config / actions and must reproduce the whole trajectory FROM THE SEED (same draw order).
"""
from __future__ import annotations

import copy
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/server"


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("ray"); mod("ray.rllib"); mod("ray.rllib.env")
    mod("ray.rllib.env.multi_agent_env", MultiAgentEnv=object)
    mod("matplotlib"); mod("matplotlib.pyplot")
    mod("perlin_noise", PerlinNoise=object)
    mod("wandb")
    mod("v0.wandb_setup", wandb_setup=lambda *a, **k: None)


def load_v0():
    _stub_modules()
    sys.path.insert(0, REF)
    from v0.config import config_dict  # noqa: E402
    from v0.env.MA_DemandResponse import MADemandResponseEnv  # noqa: E402
    from v0.utils import normStateDict  # noqa: E402

    return config_dict, MADemandResponseEnv, normStateDict


def actions_for(t: int, n: int) -> dict:
    """Deterministic pseudo-random actions (no RNG state is consumed)."""
    return {i: bool(((t * 2654435761 + i * 40503 + (t * i) % 7) >> 3) & 1) for i in range(n)}


def set_path(d, path, v):
    keys = path.split("/")
    for k in keys[:-1]:
        d = d[k]
    d[keys[-1]] = v


CASES = [
    dict(name="v0_n12_steps_constant", n=12, T=90, seed=21, over={
        "default_env_prop/power_grid_prop/base_power_mode": "constant"}),
    dict(name="v0_n20_sinus_mixture_flags", n=20, T=80, seed=22, over={
        "default_env_prop/power_grid_prop/base_power_mode": "constant",
        "default_env_prop/power_grid_prop/signal_mode": "sinusoidals",
        "default_env_prop/reward_prop/temp_penalty_mode": "mixture",
        "default_env_prop/cluster_prop/agents_comm_mode": "closed_groups",
        "default_env_prop/cluster_prop/nb_agents_comm": 4,
        "default_env_prop/state_properties/thermal": True,
        "default_env_prop/state_properties/hvac": True,
        "default_env_prop/state_properties/solar_gain": True,
        "default_env_prop/message_properties/thermal": True,
        "default_env_prop/message_properties/hvac": True,
        "default_env_prop/start_datetime_mode": "random"}),
    dict(name="v0_n9_random_fixed_flat", n=9, T=60, seed=23, over={
        "default_env_prop/power_grid_prop/base_power_mode": "constant",
        "default_env_prop/power_grid_prop/signal_mode": "flat",
        "default_env_prop/cluster_prop/agents_comm_mode": "random_fixed",
        "default_env_prop/cluster_prop/nb_agents_comm": 3,
        "default_env_prop/power_grid_prop/artificial_signal_ratio_range": 2}),
    # per-HVAC lock-out durations: lockout_duration 24 s + randint(-12, 12) (MA_DemandResponse.py:397-403)
    dict(name="v0_n14_lockout_noise", n=14, T=70, seed=24, over={
        "default_env_prop/power_grid_prop/base_power_mode": "constant",
        "default_env_prop/power_grid_prop/signal_mode": "sinusoidals",
        "default_hvac_prop/lockout_duration": 24,
        "default_hvac_prop/lockout_noise": 12}),
]

SCALARS = ["house_temp", "house_mass_temp", "hvac_turned_on", "hvac_seconds_since_off", "hvac_lockout",
           "house_target_temp", "house_Ua", "house_Cm", "house_Ca", "house_Hm", "hvac_cooling_capacity",
           "hvac_lockout_duration", "house_solar_gain", "OD_temp", "reg_signal", "cluster_hvac_power"]


def main():
    config_dict, Env, norm = load_v0()
    for case in CASES:
        cfg = copy.deepcopy(config_dict)
        cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = case["n"]
        for path, v in case["over"].items():
            set_path(cfg, path, v)
        random.seed(case["seed"])
        np.random.seed(case["seed"])
        env = Env(cfg, test=False)
        obs = env.reset()
        n, T = case["n"], case["T"]
        log = {k: np.zeros((T + 1, n)) for k in SCALARS}
        rewards = np.zeros((T, n))
        power_info = np.zeros(T)
        epochs = np.zeros(T + 1)
        msg_ids = np.array([env.cluster.agent_communicators[i] for i in range(n)])
        vecs = []

        def record(t, obs):
            for k in SCALARS:
                log[k][t] = [float(obs[i][k]) for i in range(n)]
            epochs[t] = obs[0]["datetime"].timestamp()
            vecs.append(np.stack([np.asarray(norm(obs[i], cfg), dtype=np.float64) for i in range(n)]))

        record(0, obs)
        for t in range(T):
            obs, rew, dones, info = env.step(actions_for(t, n))
            assert not any(dones.values())
            rewards[t] = [rew[i] for i in range(n)]
            power_info[t] = info["cluster_hvac_power"]
            record(t + 1, obs)
        out = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(out, n=n, T=T, seed=case["seed"], config_json=json.dumps(cfg, default=str), rewards=rewards,
                            power_info=power_info, epochs=epochs, msg_ids=msg_ids, vectors=np.stack(vecs),
                            **{"obs_" + k: v for k, v in log.items()})
        print("wrote", out, "final mean temp", log["house_temp"][-1].mean())


if __name__ == "__main__":
    main()
