"""The drop-in ``Environment`` (dict API) against the golden trajectories: same ``random.seed``,
same draw order as the reference => same reset state and same trajectory."""
import copy
import json
import random
import sys

import numpy as np
import pytest

from golden_util import GoldenCase, case_names

pytestmark = pytest.mark.gpu


def _install_perlin_stub():
    from oracle import refenv

    refenv._install_stubs()  # the same deterministic stand-in the golden files were recorded with


NAMES = [n for n in case_names()]


@pytest.mark.parametrize("name", NAMES)
def test_dropin_environment_reproduces_reference_from_seed(name):
    _install_perlin_stub()
    from marl_demandresponse_b200.environment import Environment, norm_state_dict
    from oracle.np_oracle import deadband_bangbang, greedy_myopic

    case = GoldenCase(name)
    z, N = case.z, case.N
    random.seed(case.meta["seed"])
    env = Environment(case.env_prop, interp_table=case.table())
    random.seed(case.meta["seed"])
    obs = env.reset()
    # reset state == the reference's reset state (it consumed the same Python `random` stream)
    st0 = case.state0
    snap = env._snap
    for k in ("target", "cap", "t_air", "t_mass"):
        np.testing.assert_allclose(snap[k], st0[k], rtol=0, atol=1e-12, err_msg=k)
    assert int(snap["epoch"][0]) == int(st0["epoch"][0])
    np.testing.assert_allclose(snap["od_temp"], st0["od_temp"], atol=1e-12)
    np.testing.assert_allclose(snap["signal"], st0["signal"], rtol=1e-12)
    v0 = np.array(norm_state_dict(obs, env.init_props))
    np.testing.assert_allclose(v0, z["obs"][0], rtol=0, atol=1e-9)
    assert list(obs[0].keys()) == list(case.last_obs_house0().keys())
    rng = np.random.default_rng(case.meta["seed"])
    hv = case.env_prop["cluster_prop"]["house_prop"]["hvac_prop"]
    for t in range(case.T):
        if case.meta["policy"] == "random":
            a = rng.random(N) < 0.5
        elif case.meta["policy"] == "bangbang":
            a = deadband_bangbang(np.array([obs[i]["indoor_temp"] for i in range(N)]),
                                  np.array([obs[i]["target_temp"] for i in range(N)]),
                                  np.array([obs[i]["deadband"] for i in range(N)]),
                                  np.array([obs[i]["turned_on"] for i in range(N)]))
        else:
            a = greedy_myopic(np.array([obs[i]["indoor_temp"] for i in range(N)]),
                              np.array([obs[i]["target_temp"] for i in range(N)]),
                              np.array([obs[i]["cooling_capacity"] for i in range(N)]), hv["cop"],
                              np.array([obs[i]["lockout"] for i in range(N)]), obs[0]["reg_signal"])
        assert np.array_equal(np.asarray(a, dtype=np.uint8), z["actions"][t]), f"closed-loop action diverged at step {t}"
        obs, rew = env.step({i: bool(a[i]) for i in range(N)})
        assert [obs[i]["turned_on"] for i in range(N)] == [bool(x) for x in z["on"][t]]
        assert [obs[i]["lockout"] for i in range(N)] == [bool(x) for x in z["lockout"][t]]
        assert [obs[i]["seconds_since_off"] for i in range(N)] == [int(x) for x in z["sso"][t]]
        np.testing.assert_allclose([obs[i]["indoor_temp"] for i in range(N)], z["t_air"][t], rtol=0, atol=3e-10)
        np.testing.assert_allclose([rew[i] for i in range(N)], z["rewards"][t], rtol=1e-9, atol=3e-9)
        np.testing.assert_allclose(obs[0]["reg_signal"], z["signal"][t], rtol=1e-12)
        assert env.date_time.isoformat() == obs[0]["datetime"].isoformat()
    # last dict observation of house 0: same keys / order / values / message list as the reference
    want = case.last_obs_house0()
    got = obs[0]
    assert list(got.keys()) == list(want.keys())
    assert got["datetime"].isoformat() == want["datetime"]
    assert len(got["message"]) == len(want["message"])
    for mg, mw in zip(got["message"], want["message"]):
        assert list(mg.keys()) == list(mw.keys())
        for kk, vv in mw.items():
            assert abs(float(mg[kk]) - float(vv)) <= 1e-9 * max(1.0, abs(float(vv))), kk
    for k, v in want.items():
        if k in ("datetime", "message"):
            continue
        assert abs(float(got[k]) - float(v)) <= 1e-9 * max(1.0, abs(float(v))), k
    vT = np.array(norm_state_dict(obs, env.init_props))
    if case.obs_stride == 1 or case.T % case.obs_stride == 0:
        np.testing.assert_allclose(vT, z["obs"][case.T // case.obs_stride], rtol=0, atol=1e-9)


def test_dropin_deepcopy_and_attributes():
    _install_perlin_stub()
    from marl_demandresponse_b200.environment import Environment

    case = GoldenCase("c1_default_n10_bangbang")
    random.seed(1)
    env = Environment(case.env_prop)
    env.step({i: True for i in range(case.N)})
    twin = copy.deepcopy(env)
    state = random.getstate()
    o1, r1 = env.step({0: True})
    random.setstate(state)
    o2, r2 = twin.step({0: True})
    assert r1 == r2 and o1[3]["indoor_temp"] == o2[3]["indoor_temp"]
    b = env.cluster.buildings[2]
    assert b.indoor_temp == o1[2]["indoor_temp"] and b.hvac.turned_on == o1[2]["turned_on"]
    assert env.cluster.max_power == case.N * 6000.0
    assert env.power_grid.current_signal == o1[0]["reg_signal"]
    assert env.current_od_temp == o1[0]["OD_temp"]
