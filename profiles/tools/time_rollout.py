#!/usr/bin/env python
"""Device time of the on-device rollout transition (actor + draw + env step) next to its two halves.

    python profiles/tools/time_rollout.py R N layout
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from bench import env_prop_for
from marl_demandresponse_b200 import BatchedEnv

R, N, layout = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
env = BatchedEnv(env_prop_for(N), R, precision="f32", obs_layout=layout, policy="external", noise="philox", seed=1)
env.reset()
D = env.sim.D
torch.manual_seed(0)
fc = torch.nn.ModuleList([torch.nn.Linear(D, 100), torch.nn.Linear(100, 100), torch.nn.Linear(100, 2)]).cuda()
w = BatchedEnv.actor_weights(fc)
acts = (torch.rand((R, N), device="cuda") < 0.5).to(torch.uint8)


def timed(fn, k=200):
    """CUDA-graph replay when the work is short enough to be bound by the Python launch rate."""
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / k


def torch_policy():
    with torch.no_grad():
        x = env.obs.reshape(-1, D)
        p = torch.softmax(fc[2](torch.relu(fc[1](torch.relu(fc[0](x))))), dim=1)
        a = torch.multinomial(p, 1)
    return a


flops = 2.0 * R * N * (D * 100 + 100 * 100 + 100 * 2)
t_env = timed(lambda: env.step(acts))
t_pol = timed(lambda: env.policy_step(w))
t_all = timed(lambda: env.rollout_step(w))
t_torch = timed(torch_policy, 50)
print(f"R={R} N={N} {layout} D={D}: env step {t_env:.1f} us | actor+draw {t_pol:.1f} us ({flops / t_pol / 1e6:.1f} TFLOP/s, "
      f"{R * N * (4 * D + 5) / t_pol / 1e3:.0f} GB/s) | rollout transition {t_all:.1f} us = {R * N / t_all / 1e3:.2f} G agent-steps/s | "
      f"torch fp32 actor + multinomial {t_torch:.1f} us")
