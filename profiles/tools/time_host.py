#!/usr/bin/env python
"""Wall-clock breakdown of the host-buffer step (drsim_step_host) against the device-resident step.

    python profiles/tools/time_host.py R N layout
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from bench import env_prop_for
from marl_demandresponse_b200 import BatchedEnv

R, N, layout = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
env = BatchedEnv(env_prop_for(N), R, precision="f32", obs_layout=layout, policy="external", noise="philox", seed=1)
env.reset()
acts = [(torch.rand((R, N), device="cuda") < 0.5).to(torch.uint8) for _ in range(4)]
pinned = [a.cpu().pin_memory() for a in acts]
pageable = [a.cpu().numpy().copy() for a in acts]
out = torch.zeros((R, 4), dtype=torch.float64).pin_memory()
K = 300


def wall(fn):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e6


def dev_sync(i):
    env.step(acts[i % 4])
    torch.cuda.synchronize()


print(f"R={R} N={N} {layout}: us per step (wall clock)")
print(f"  device step, no sync          {wall(lambda i: env.step(acts[i % 4])):8.1f}")
print(f"  device step + synchronize     {wall(dev_sync):8.1f}")
print(f"  step_host pinned (zero-copy)  {wall(lambda i: env.step_host(pinned[i % 4], out)):8.1f}")
print(f"  step_host pageable (copies)   {wall(lambda i: env.step_host(pageable[i % 4], out)):8.1f}")
print(f"  step_host pinned, no results  {wall(lambda i: env.sim.step_host(pinned[i % 4], env_out=out)):8.1f}")
