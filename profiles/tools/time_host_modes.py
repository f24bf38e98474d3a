"""e2e host-buffer step (drsim_step_host, pinned actions in, per-replica results out, sync) in both
action-transfer modes.  usage: time_host_modes.py R N layout"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
R, N, layout = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
for mode in ("zerocopy", "dma"):
    os.environ["DRSIM_HOST_ACTIONS"] = mode
    from marl_demandresponse_b200 import BatchedEnv

    prop = {"cluster_prop": {"nb_agents": N}, "start_datetime": "2021-06-15T12:00:00", "start_datetime_mode": "fixed"}
    env = BatchedEnv(prop, R, obs_layout=layout, noise="philox", seed=1)
    env.reset()
    acts = [(torch.rand((R, N)) < 0.5).to(torch.uint8).pin_memory() for _ in range(4)]
    out = torch.zeros((R, 4), dtype=torch.float64).pin_memory()
    for i in range(300):
        env.step_host(acts[i % 4], out)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 500
        for i in range(n):
            env.step_host(acts[i % 4], out)
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) / n * 1e6
        print(f"{mode:9s} R={R} N={N} {layout}: {us:7.1f} us/step  {R * N / us * 1e6:.3e} house-steps/s", flush=True)
    del env
