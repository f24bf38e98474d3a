"""The NumPy oracle against the golden trajectories recorded from the real reference."""
import numpy as np
import pytest

from golden_util import GoldenCase, case_names, replay
from oracle.np_oracle import NpOracle


@pytest.mark.parametrize("name", case_names())
def test_np_oracle_matches_reference_golden(name):
    case = GoldenCase(name)
    orc = NpOracle(case.env_prop, 1, table=case.table())
    worst = replay(case, orc, rtol=1e-12, precision="f64")
    # fp64 literal restatement: agreement is at rounding level, far inside 1e-12
    assert worst["t_air"] < 1e-11 and worst["rewards"] < 1e-11


def test_golden_cases_present():
    assert len(case_names()) >= 10


@pytest.mark.parametrize("name", [n for n in case_names() if "random" not in n.split("_")[0]])
def test_static_comm_tables_match_reference(name):
    from oracle.np_oracle import comm_table

    case = GoldenCase(name)
    if "comm_table" not in case.z:
        pytest.skip("per-step tables")
    if case.env_prop["cluster_prop"]["agents_comm_prop"]["mode"] == "random_fixed":
        pytest.skip("drawn from Python random at reset")
    got = comm_table(case.N, case.env_prop["cluster_prop"]["agents_comm_prop"])
    assert np.array_equal(got, case.z["comm_table"])


def test_controllers_reproduce_golden_actions():
    from oracle.np_oracle import deadband_bangbang, greedy_myopic

    for name, pol in (("c1_default_n10_bangbang", "bb"), ("greedy_n40", "greedy")):
        case = GoldenCase(name)
        z = case.z
        hv = case.env_prop["cluster_prop"]["house_prop"]["hvac_prop"]
        st = case.state0
        t_air, on, lock, sig = st["t_air"][0], st["on"][0], st["lockout"][0], st["signal"][0]
        for t in range(case.T):
            if pol == "bb":
                a = deadband_bangbang(t_air, st["target"][0], case.env_prop["cluster_prop"]["house_prop"]["deadband"], on)
            else:
                a = greedy_myopic(t_air, st["target"][0], st["cap"][0], hv["cop"], lock, sig)
            assert np.array_equal(a.astype(np.uint8), z["actions"][t]), (name, t)
            t_air, on, lock, sig = z["t_air"][t], z["on"][t], z["lockout"][t], z["signal"][t]


@pytest.mark.parametrize("name", [n for n in case_names() if "allflags" not in n])
def test_scalar_port_matches_reference_golden(name):
    """The per-house pure-Python port (the timed CPU baseline) against the same fixtures."""
    from oracle.scalar_port import ScalarEnv

    case = GoldenCase(name)
    env = ScalarEnv(case.env_prop, table=case.table())

    class Adapter:
        def set_state(self, st):
            env.set_state(st)

        def get_state(self):
            return env.get_state()

        def step(self, a, od, perlin, ids):
            t = self.t
            self.t += 1
            return env.step(a, od, perlin, ids, comm=case.comm(t + 1))

    ad = Adapter()
    ad.t = 0
    replay(case, ad, rtol=1e-12, check_obs=False, precision="f64")
    # dict observation of house 0 after the last step equals the reference's (keys, order, values)
    got, want = env.last_obs[0], case.last_obs_house0()
    assert list(got.keys()) == list(want.keys())
    for k, v in want.items():
        if k == "datetime":
            assert got[k].isoformat() == v
        elif k == "message":
            assert len(got[k]) == len(v)
            for mg, mw in zip(got[k], v):
                assert set(mw.keys()) <= set(mg.keys()) | {"cop", "latent_cooling_fraction", "cooling_capacity", "Ua", "Ca", "Cm", "Hm"}
                for kk in ("seconds_since_off", "curr_consumption", "max_consumption", "current_temp_diff_to_target"):
                    assert abs(float(mg[kk]) - float(mw[kk])) <= 1e-9 * max(1.0, abs(float(mw[kk])))
        else:
            assert abs(float(got[k]) - float(v)) <= 1e-9 * max(1.0, abs(float(v))), k
