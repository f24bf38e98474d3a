#!/usr/bin/env python
"""Host-side launch floor of the house-sharded step: one rank, a cluster small enough that the GPU work is
negligible -- what remains is Python + ctypes + four kernel launches per step.

    python profiles/tools/time_sharded_host.py [N]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from bench import env_prop_for
from marl_demandresponse_b200.sharded import ShardedClusterEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
env = ShardedClusterEnv(env_prop_for(N), 1, rank=0, world=1, device=0, obs_layout="tarmac", noise="philox", seed=1, exchange="none")
env.reset()
for _ in range(200):
    env.step(None)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 3000
for _ in range(K):
    env.step(None)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"N={N}: host issue {1e6 * (t1 - t0) / K:.1f} us/step, with final sync {1e6 * (t2 - t0) / K:.1f} us/step")
