import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench import env_prop_for
from marl_demandresponse_b200 import BatchedEnv
from marl_demandresponse_b200.batched import synthetic_state
R, N, layout = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
prop = env_prop_for(N)
st = synthetic_state(prop, R, seed=3)
acts = (np.random.default_rng(2).random((6, R, N)) < 0.5).astype(np.uint8)
out = {}
for path in ("fused", "split"):
    env = BatchedEnv(prop, R, obs_layout=layout, noise="philox", seed=9, path=path)
    print(path, env.sim.fused_info(), flush=True)
    env.reset(copy.deepcopy(st))
    for t in range(6):
        env.step(torch.as_tensor(acts[t], device="cuda"))
        torch.cuda.synchronize()
    out[path] = {k: env.state[k].clone() for k in ("sso", "flags", "dt_air", "reward", "obs", "signal", "power")}
for k in out["fused"]:
    a, b = out["fused"][k].double(), out["split"][k].double()
    print(k, float((a - b).abs().max()))
