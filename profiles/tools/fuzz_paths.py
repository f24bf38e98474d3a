#!/usr/bin/env python
"""Randomised cross-check of the fused kernels (whatever variant / tile size drsim_create picks) against the
general path over many (R, N, layout, flags) combinations: discrete state and env scalars bit-exact,
continuous values to fp32 rounding.  Prints one line per configuration; exits non-zero on a mismatch.

    python profiles/tools/fuzz_paths.py [n_configs] [seed] [f32|f64]
"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from marl_demandresponse_b200 import BatchedEnv
from marl_demandresponse_b200.batched import synthetic_state

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
precision = sys.argv[3] if len(sys.argv) > 3 else "f32"
bad = 0
for it in range(n_cfg):
    N = int(rng.choice([1, 2, 3, 5, 9, 10, 12, 33, 64, 100, 127, 128, 250, 333, 512, 1000, 1023, 1024]))
    R = int(rng.integers(1, max(2, min(6000, 600000 // N))))
    layout = str(rng.choice(["tarmac", "hand_engineered", "none"]))
    pen = str(rng.choice(["individual_L2", "common_L2", "common_max_error", "mixture"], p=[0.6, 0.15, 0.1, 0.15]))
    sig = str(rng.choice(["perlin", "sinusoidals", "regular_steps", "flat"]))
    prop = {"start_datetime": "2021-06-15T11:58:20", "start_datetime_mode": "fixed", "time_step": 4.0,
            "cluster_prop": {"nb_agents": N, "house_prop": {"target_temp": 19.0},
                             "agents_comm_prop": {"max_nb_agents_communication": int(rng.choice([0, 1, 4, 10, 11]))}},
            "reward_prop": {"penalty_props": {"mode": pen}},
            "state_prop": {"solar_gain": bool(rng.random() < 0.2), "thermal": bool(rng.random() < 0.2), "hvac": bool(rng.random() < 0.2)},
            "power_grid_prop": {"signal_properties": {"mode": sig}}}
    T = 5
    st = synthetic_state(prop, R, seed=it)
    acts = (rng.random((T, R, N)) < 0.5).astype(np.uint8)
    out = {}
    info = None
    try:
        for path in ("auto", "split"):
            env = BatchedEnv(prop, R, obs_layout=layout, noise="philox", seed=7, path=path, precision=precision)
            if path == "auto":
                info = env.sim.fused_info()
            env.reset(copy.deepcopy(st))
            for t in range(T):
                env.step(torch.as_tensor(acts[t], device="cuda"))
            torch.cuda.synchronize()
            keys = ["signal", "od_temp", "epoch", "sso", "flags", "power", "dt_air" if precision == "f32" else "t_air", "reward", "metrics"] + (["obs"] if layout != "none" else [])
            out[path] = {k: env.state[k].clone() for k in keys}
            del env
        msg = "ok"
        for k in ("signal", "od_temp", "epoch", "sso", "flags"):
            if not torch.equal(out["auto"][k], out["split"][k]):
                msg = f"MISMATCH {k}"
        for k in out["auto"]:
            if k in ("signal", "od_temp", "epoch", "sso", "flags"):
                continue
            a, b = out["auto"][k].double(), out["split"][k].double()
            if not torch.allclose(a, b, rtol=3e-6, atol=3e-6):
                msg = f"MISMATCH {k} max {float((a - b).abs().max()):.3e}"
    except Exception as e:  # noqa: BLE001
        msg = f"ERROR {type(e).__name__}: {e}"
    if msg != "ok":
        bad += 1
    print(f"{it:3d} R={R:5d} N={N:4d} {layout:15s} {pen:16s} {sig:13s} {info} -> {msg}", flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)
