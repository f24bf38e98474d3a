"""Per-phase timing of k_shard on a house-sharded cluster across ranks (torchrun): rank 0 prints its CTAs' stamps.

    DRSIM_SHARD_DBG=1 python -m torch.distributed.run --nproc-per-node 8 ... profiles/tools/shard_phases_multi.py [n_houses]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("DRSIM_SHARD_DBG", "1")
from marl_demandresponse_b200.sharded import ShardedClusterEnv  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
prop = {"start_datetime": "2021-06-15T12:00:00", "start_datetime_mode": "fixed", "time_step": 4.0,
        "cluster_prop": {"nb_agents": n, "house_prop": {"target_temp": 19.0}}}
env = ShardedClusterEnv(prop, 1, rank=rank, world=world, device=local, obs_layout="tarmac", noise="philox", seed=1234, exchange="peer")
env.reset()
tape = (torch.rand((4, 1, env.hi - env.lo), device="cuda") < 0.5).to(torch.uint8)
env.run(200, tape, rotate=True)
torch.cuda.synchronize()
dist.barrier()
env.run(200, tape, rotate=True)
torch.cuda.synchronize()
L = env.sim._L
L.drsim_debug_shard_times.restype = C.c_int
L.drsim_debug_shard_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
buf = np.zeros((1024, 16), dtype=np.uint64)
g = L.drsim_debug_shard_times(env.sim._h, buf.ctypes.data_as(C.c_void_p), 1024)
t = buf[:g].astype(np.int64)
for r in range(world):
    dist.barrier()
    if r == rank and rank in (0, world - 1):
        t0 = t[:, 0].min()
        w = t[1:]   # CTA 0 is the dedicated reducer
        f = lambda k: "%.2f / %.2f / %.2f" % tuple(((w[:, k] - t0) / 1e3)[[0, 0, 0]] * 0 + np.array([((w[:, k] - t0) / 1e3).min(), np.median((w[:, k] - t0) / 1e3), ((w[:, k] - t0) / 1e3).max()]))
        print(f"rank {rank}: {g} CTAs; us from the first CTA's start (min / median / max over the working CTAs)")
        for k, nm in ((0, "start"), (1, "dep-wait done"), (14, "pre-pass + early reduce done"), (2, "phase 1 done"), (3, "late partials out"),
                      (4, "cluster power seen"), (5, "end")):
            print(f"   {nm:30s} {f(k)}")
        r0 = (t[0] - t0) / 1e3
        print(f"   reducer CTA: early done {r0[14]:.2f}, late collected {r0[6]:.2f}, late rows combined {r0[7]:.2f}, stored {r0[3]:.2f}", flush=True)
dist.destroy_process_group()
