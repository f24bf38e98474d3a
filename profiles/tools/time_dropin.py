"""Wall time per step of the drop-in ``Environment`` (dict API, fp64 build), the CPU port next to it.

    python profiles/tools/time_dropin.py [n_houses ...]
"""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from marl_demandresponse_b200 import Environment  # noqa: E402
from marl_demandresponse_b200.environment import norm_state_dict  # noqa: E402

for n in [int(x) for x in sys.argv[1:]] or [10, 1000]:
    prop = {"start_datetime": "2021-06-15T12:00:00", "start_datetime_mode": "fixed", "time_step": 4.0,
            "cluster_prop": {"nb_agents": n, "house_prop": {"target_temp": 19.0}},
            "power_grid_prop": {"signal_properties": {"mode": "sinusoidals"}}}
    random.seed(4)
    env = Environment(prop)
    obs = env.reset()
    rng = np.random.default_rng(0)
    acts = [dict(enumerate((rng.random(n) < 0.5).tolist())) for _ in range(8)]
    T = 300 if n <= 100 else 100
    res = {}
    for mode in ("vectors", "all_dicts", "bangbang_dicts"):
        for t in range(20):
            env.step(acts[t % 8])
        t0 = time.perf_counter()
        for t in range(T):
            if mode == "bangbang_dicts":   # a per-house Python controller reading its own dict (DeadbandBangBang)
                a = {i: obs[i]["indoor_temp"] > obs[i]["target_temp"] for i in range(n)}
            else:
                a = acts[t % 8]
            obs, rew = env.step(a)
            if mode == "vectors":
                v = norm_state_dict(obs, env.init_props)
            elif mode == "all_dicts":
                obs.materialize()
        res[mode] = (time.perf_counter() - t0) / T * 1e6
    print(f"N={n}: " + ", ".join(f"{k} {v:.0f} us/step" for k, v in res.items()), flush=True)
