#!/usr/bin/env python
"""Micro-timing of BatchedEnv.step for a few layouts (CUDA events, 200 steps after 20 warm-up).

    python profiles/tools/time_step.py R N layout [layout ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from bench import env_prop_for, algorithmic_bytes_per_house_step
from marl_demandresponse_b200 import BatchedEnv

R, N = int(sys.argv[1]), int(sys.argv[2])
for layout in sys.argv[3:]:
    env = BatchedEnv(env_prop_for(N), R, precision="f32", obs_layout=layout, policy="external", noise="philox", seed=1)
    env.reset()
    acts = [(torch.rand((R, N), device="cuda") < 0.5).to(torch.uint8) for _ in range(4)]
    for i in range(20):
        env.step(acts[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        env.step(acts[i % 4])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    b = algorithmic_bytes_per_house_step(4, env.sim.D)
    print(f"R={R} N={N} {layout:16s} D={env.sim.D:3d} {us:7.2f} us/step  {R * N * b / us / 1e3:7.0f} GB/s ({b} B/house)")
    del env
