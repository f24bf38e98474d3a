"""Legacy (v0) host-side pieces against the RUNNING legacy environment of the reference (authoring
container only; skipped where ``/root/reference`` is absent): ``norm_state_dict_v0`` on the observation
dicts the real ``MADemandResponseEnv`` produces over random configurations (all state / message flags,
communication modes, signal modes), bit for bit, and ``props_from_v0`` on the same configurations."""
import copy
import os
import random
import sys

import numpy as np
import pytest

from oracle import refenv

pytestmark = pytest.mark.skipif(not refenv.available(), reason="needs the reference checkout (/root/reference)")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def legacy():
    import make_golden_v0 as g

    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "perlin_noise")}
    try:
        return g, g.load_v0()
    finally:
        # load_v0 registers its own stand-ins for modules other tests stub differently: restore them
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)


@pytest.mark.parametrize("seed", range(10))
def test_legacy_norm_vector_and_config_translation(seed, legacy):
    from marl_demandresponse_b200.v0 import norm_state_dict_v0, props_from_v0

    g, (config_dict, Env, norm_ref) = legacy
    rng = random.Random(300 + seed)
    cfg = copy.deepcopy(config_dict)
    env_p = cfg["default_env_prop"]
    n = rng.choice([4, 9, 12, 20])
    env_p["cluster_prop"]["nb_agents"] = n
    env_p["cluster_prop"]["agents_comm_mode"] = rng.choice(["neighbours", "closed_groups", "random_fixed"])
    env_p["cluster_prop"]["nb_agents_comm"] = rng.choice([2, 3])
    env_p["power_grid_prop"]["base_power_mode"] = "constant"
    env_p["power_grid_prop"]["signal_mode"] = rng.choice(["flat", "sinusoidals", "regular_steps"])
    env_p["reward_prop"]["temp_penalty_mode"] = rng.choice(["individual_L2", "common_L2", "common_max", "mixture"])
    for k in ("thermal", "hvac", "solar_gain", "hour", "day"):
        if k in env_p["state_properties"]:
            env_p["state_properties"][k] = rng.random() < 0.5
    for k in ("thermal", "hvac"):
        env_p["message_properties"][k] = rng.random() < 0.5
    state = random.getstate()
    try:
        random.seed(40 + seed)
        np.random.seed(40 + seed)
        try:
            env = Env(cfg, test=False)
            obs = env.reset()
        except (IndexError, ValueError, KeyError) as e:
            pytest.skip(f"the legacy reference itself fails on this configuration: {type(e).__name__}: {e}")
        for t in range(12):
            for i in range(n):
                want = np.asarray(norm_ref(obs[i], cfg), dtype=np.float64)
                got = np.asarray(norm_state_dict_v0(obs[i], cfg), dtype=np.float64)
                assert got.shape == want.shape, (t, i)
                assert np.array_equal(got, want), (t, i, np.abs(got - want).max())
            obs, _, _, _ = env.step(g.actions_for(t, n))
    finally:
        random.setstate(state)
    p = props_from_v0(cfg)
    assert p.cluster_prop.nb_agents == n
    assert p.cluster_prop.agents_comm_prop.mode == env_p["cluster_prop"]["agents_comm_mode"]
    assert p.cluster_prop.agents_comm_prop.max_nb_agents_communication == env_p["cluster_prop"]["nb_agents_comm"]
    assert p.power_grid_prop.signal_properties.mode == env_p["power_grid_prop"]["signal_mode"]
    assert p.state_prop.thermal == env_p["state_properties"]["thermal"] and p.state_prop.hvac == env_p["state_properties"]["hvac"]
    assert p.cluster_prop.message_prop.thermal == env_p["message_properties"]["thermal"]
