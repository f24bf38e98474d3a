"""Where a drop-in ``Environment.step`` spends its wall time: the C call (wrapper included), the rest of ``step``, ``norm_state_dict``.

    python profiles/tools/prof_dropin.py
"""
import os, random, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from marl_demandresponse_b200 import Environment
from marl_demandresponse_b200.environment import norm_state_dict
for n in (10, 100, 1000):
    prop = {"start_datetime": "2021-06-15T12:00:00", "start_datetime_mode": "fixed", "time_step": 4.0,
            "cluster_prop": {"nb_agents": n, "house_prop": {"target_temp": 19.0}},
            "power_grid_prop": {"signal_properties": {"mode": "sinusoidals"}}}
    random.seed(4)
    env = Environment(prop)
    obs = env.reset()
    rng = np.random.default_rng(0)
    acts = [dict(enumerate((rng.random(n) < 0.5).tolist())) for _ in range(8)]
    for t in range(30):
        env.step(acts[t % 8])
    T = 300
    # C call alone
    a = np.zeros((1, n), dtype=np.uint8); od = np.zeros(1); 
    t0 = time.perf_counter()
    for t in range(T):
        env._sim.step_host_snapshot(a, od, None, None)
    c_call = (time.perf_counter() - t0) / T * 1e6
    t0 = time.perf_counter()
    for t in range(T):
        o, r = env.step(acts[t % 8])
    step = (time.perf_counter() - t0) / T * 1e6
    t0 = time.perf_counter()
    for t in range(T):
        o, r = env.step(acts[t % 8]); v = norm_state_dict(o, env.init_props)
    full = (time.perf_counter() - t0) / T * 1e6
    print(f"N={n}: C call (python wrapper incl.) {c_call:.0f} us, env.step {step:.0f} us, step + norm_state_dict {full:.0f} us", flush=True)
