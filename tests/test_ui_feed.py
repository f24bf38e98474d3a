"""The UI feed (``marl_demandresponse_b200.ui_feed.ClientFeed``) against payloads recorded from the
reference's own ``ClientManagerService`` (``tests/golden/make_golden_ui.py``).

CPU part: fed the recorded observation dicts, the feed reproduces the reference's ``dataChange`` /
``houseChange`` payloads and graph series.  GPU part: the drop-in ``Environment`` replays the same
``random.seed`` trajectory and the feed is driven from the environment's tensors; a ``BatchedEnv``
replica goes through ``drsim_cluster_summary``."""
import asyncio
import json
import os
import random

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ui_feed_n10", "ui_feed_n33_sinus"]
ROUNDED = {"Outdoor temperature", "Average indoor temperature", "Average temperature difference", "Mass temperature",
           "Target temperature"}


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def check_description(got, want, rtol):
    assert list(got.keys()) == list(want.keys())
    for k, w in want.items():
        g = got[k]
        if k in ("Number of HVAC", "Number of locked HVAC"):
            assert g == w, k
        elif k in ROUNDED:
            # two-decimal strings: a value within `rtol` of a rounding boundary may land on the other side
            assert abs(float(g) - float(w)) <= 0.01 + 1e-9, (k, g, w)
            assert len(g.split(".")[-1]) <= 2
        else:
            assert abs(float(g) - float(w)) <= rtol * max(1.0, abs(float(w))), (k, g, w)


def check_houses(got, want, atol):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert list(g.keys()) == list(w.keys()), (g, w)
        assert g["id"] == w["id"] and g["hvacStatus"] == w["hvacStatus"]
        if "secondsSinceOff" in w:
            assert g["secondsSinceOff"] == w["secondsSinceOff"]
        for k in ("indoorTemp", "targetTemp", "tempDifference"):
            assert abs(g[k] - w[k]) <= atol, (k, g[k], w[k])


class Socket:
    def __init__(self):
        self.sent = []

    async def emit(self, endpoint, data):
        self.sent.append(endpoint)


@pytest.mark.parametrize("name", CASES)
def test_feed_from_recorded_observations_matches_reference(name):
    from marl_demandresponse_b200.ui_feed import DESCRIPTION_KEYS, ClientFeed

    z = load(name)
    assert DESCRIPTION_KEYS == z["description_keys"]
    sock = Socket()
    feed = ClientFeed(sock)
    feed.initialize_data(True)
    for t, st in enumerate(z["steps"]):
        obs = {int(i): o for i, o in st["obs"].items()}
        asyncio.run(feed.update_data(obs, t))
        assert sock.sent[-2:] == st["emitted"]
        check_description(feed.description[t], st["description"], 1e-12)
        check_houses(feed.houses_data[t], st["houses"], 0.0)
    for k, want in z["series"].items():
        np.testing.assert_allclose(getattr(feed, k), want, rtol=1e-13, atol=1e-13, err_msg=k)
    # without an interface nothing is emitted (client_manager_service.py:266-267)
    quiet = ClientFeed(sock)
    quiet.initialize_data(False)
    n = len(sock.sent)
    asyncio.run(quiet.update_data({int(i): o for i, o in z["steps"][0]["obs"].items()}, 0))
    assert len(sock.sent) == n


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_feed_from_dropin_environment_matches_reference(name):
    from oracle import refenv
    from oracle.np_oracle import deadband_bangbang

    refenv._install_stubs()  # the deterministic perlin stand-in the recording used
    from marl_demandresponse_b200 import ClientFeed, Environment

    z = load(name)
    N = z["env_prop"]["cluster_prop"]["nb_agents"]
    random.seed(z["seed"])
    env = Environment(z["env_prop"])
    random.seed(z["seed"])
    obs = env.reset()
    feed, feed_dict = ClientFeed(), ClientFeed()
    for t, st in enumerate(z["steps"]):
        a = deadband_bangbang(np.array([obs[i]["indoor_temp"] for i in range(N)]),
                              np.array([obs[i]["target_temp"] for i in range(N)]),
                              np.array([obs[i]["deadband"] for i in range(N)]),
                              np.array([obs[i]["turned_on"] for i in range(N)]))
        assert [int(x) for x in a] == st["actions"], f"closed-loop action diverged at step {t}"
        obs, _ = env.step({i: bool(a[i]) for i in range(N)})
        houses, desc = feed.update(env, t)            # from the environment's tensors
        check_description(desc, st["description"], 1e-9)
        check_houses(houses, st["houses"], 1e-9)
        h2, d2 = feed_dict.update(obs, t)             # from the dicts the drop-in hands out: same thing
        assert d2 == desc and h2 == houses
    for k, want in z["series"].items():
        np.testing.assert_allclose(getattr(feed, k), want, rtol=1e-9, atol=1e-9, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_cluster_summary_and_batched_feed(precision):
    import torch

    from marl_demandresponse_b200 import BatchedEnv, ClientFeed
    from marl_demandresponse_b200.batched import synthetic_state

    R, N = 7, 1003
    prop = {"cluster_prop": {"nb_agents": N}, "power_grid_prop": {"signal_properties": {"mode": "sinusoidals"}}}
    env = BatchedEnv(prop, R, precision=precision, obs_layout="tarmac", noise="philox", seed=5)
    env.reset(synthetic_state(prop, R, seed=8))
    rng = np.random.default_rng(0)
    feed = ClientFeed()
    tol = 1e-4 if precision == "f32" else 1e-10
    for t in range(12):
        env.step(torch.as_tensor((rng.random((R, N)) < 0.4).astype(np.uint8), device="cuda"))
        s = env.get_state()
        summ = env.sim.cluster_summary().cpu().numpy()
        again = env.sim.cluster_summary().cpu().numpy()
        assert np.array_equal(summ, again)              # fixed summation order
        d = s["t_air"] - s["target"]
        assert np.array_equal(summ[:, 0], s["lockout"].sum(1)) and np.array_equal(summ[:, 6], s["on"].sum(1))
        assert np.all(summ[:, 7] == N)
        np.testing.assert_allclose(summ[:, 1], s["t_air"].sum(1), rtol=1e-6 if precision == "f32" else 1e-13)
        np.testing.assert_allclose(summ[:, 2], d.sum(1), rtol=0, atol=tol * N)
        np.testing.assert_allclose(summ[:, 3], np.abs(d).sum(1), rtol=0, atol=tol * N)
        np.testing.assert_allclose(summ[:, 4], s["t_mass"].sum(1), rtol=1e-6 if precision == "f32" else 1e-13)
        np.testing.assert_allclose(summ[:, 5], s["target"].sum(1), rtol=1e-6 if precision == "f32" else 1e-13)
        r = 3
        houses, desc = feed.update(env, t, replica=r)
        ref = ClientFeed()
        ref.signal, ref.consumption, ref.temp_err = feed.signal[:-1], feed.consumption[:-1], feed.temp_err[:-1]
        obs = {i: {"turned_on": bool(s["on"][r, i]), "lockout": bool(s["lockout"][r, i]), "seconds_since_off": int(s["sso"][r, i]),
                   "indoor_temp": s["t_air"][r, i], "mass_temp": s["t_mass"][r, i], "target_temp": s["target"][r, i],
                   "OD_temp": s["od_temp"][r], "reg_signal": s["signal"][r], "cluster_hvac_power": s["power"][r]} for i in range(N)}
        h_ref, d_ref = ref.update(obs, t)
        check_description(desc, d_ref, 1e-5 if precision == "f32" else 1e-11)
        check_houses(houses, h_ref, 1e-5 if precision == "f32" else 1e-11)
    assert len(feed.signal) == 12 and feed.houses_data.keys() == set(range(12))
