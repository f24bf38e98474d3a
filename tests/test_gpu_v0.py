"""Legacy (v0) gym-style adapter on the GPU against trajectories recorded from the reference's
``server/v0/env/MA_DemandResponse.py`` (``tests/golden/make_golden_v0.py``): the same Python ``random``
seed must reproduce construction, ``reset()`` and every step -- discrete state bit-exact, continuous
values to 1e-9 (fp64 build), ``normStateDict`` vectors included."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["v0_n12_steps_constant", "v0_n20_sinus_mixture_flags", "v0_n9_random_fixed_flat", "v0_n14_lockout_noise"]


def _actions_for(t, n):
    return {i: bool(((t * 2654435761 + i * 40503 + (t * i) % 7) >> 3) & 1) for i in range(n)}


@pytest.mark.parametrize("name", CASES)
def test_v0_adapter_reproduces_the_legacy_env_from_the_seed(name):
    from marl_demandresponse_b200.v0 import MADemandResponseEnv, norm_state_dict_v0

    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["config_json"]))
    n, T, seed = int(z["n"]), int(z["T"]), int(z["seed"])
    random.seed(seed)
    np.random.seed(seed)
    env = MADemandResponseEnv(cfg, test=False)
    obs = env.reset()
    assert np.array_equal(np.array([env.cluster.agent_communicators[i] for i in range(n)]), z["msg_ids"])
    exact = ("hvac_turned_on", "hvac_seconds_since_off", "hvac_lockout", "hvac_cooling_capacity", "hvac_lockout_duration")

    def check(t, obs):
        for k in [f[4:] for f in z.files if f.startswith("obs_")]:
            got = np.array([float(obs[i][k]) for i in range(n)])
            want = z["obs_" + k][t]
            if k in exact:
                assert np.array_equal(got, want), (t, k)
            else:
                np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9, err_msg=f"{name} t={t} {k}")
        assert obs[0]["datetime"].timestamp() == z["epochs"][t]
        vec = np.stack([norm_state_dict_v0(obs[i], cfg) for i in range(n)])
        np.testing.assert_allclose(vec, z["vectors"][t], rtol=1e-9, atol=1e-9)

    check(0, obs)
    for t in range(T):
        obs, rew, dones, info = env.step(_actions_for(t, n))
        assert set(obs) == set(range(n)) and not any(dones.values())
        np.testing.assert_allclose([rew[i] for i in range(n)], z["rewards"][t], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(info["cluster_hvac_power"], z["power_info"][t], rtol=1e-12)
        check(t + 1, obs)
