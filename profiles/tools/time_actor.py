"""Actor kernel timings: single-pass TF32 against 3xTF32, on the C3 / C4 observation shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import env_prop_for
from marl_demandresponse_b200 import BatchedEnv
for name, R, N, layout in (("c3", 4096, 100, "hand_engineered"), ("c4", 2048, 1000, "tarmac")):
    env = BatchedEnv(env_prop_for(N), R, obs_layout=layout, noise="philox", seed=1)
    env.reset()
    D = env.sim.D
    torch.manual_seed(0)
    fc = torch.nn.ModuleList([torch.nn.Linear(D, 100), torch.nn.Linear(100, 100), torch.nn.Linear(100, 2)]).cuda()
    w = BatchedEnv.actor_weights(fc)
    for prec in ("tf32", "tf32x3"):
        for _ in range(5):
            env.policy_step(w, precision=prec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(100):
            env.policy_step(w, precision=prec)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10
        e0.record()
        for _ in range(100):
            env.rollout_step(w, precision=prec)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} D={D} {prec}: actor {us:.1f} us, transition {e0.elapsed_time(e1) * 10:.1f} us", flush=True)
