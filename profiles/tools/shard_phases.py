"""Per-phase timing of k_shard (csrc/drsim_shard.cuh) from the globaltimer stamps its CTAs leave when
DRSIM_SHARD_DBG=1: where one house-sharded / large-cluster step spends its time.

    DRSIM_SHARD_DBG=1 python profiles/tools/shard_phases.py [n_houses] [replicas]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("DRSIM_SHARD_DBG", "1")
from marl_demandresponse_b200.sharded import ShardedClusterEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prop = {"start_datetime": "2021-06-15T12:00:00", "start_datetime_mode": "fixed", "time_step": 4.0,
        "cluster_prop": {"nb_agents": n, "house_prop": {"target_temp": 19.0}}}
env = ShardedClusterEnv(prop, R, obs_layout="tarmac", noise="philox", seed=1234)
env.reset()
a = (torch.rand((R, env.hi - env.lo), device="cuda") < 0.5).to(torch.uint8)
for _ in range(50):
    env.step(a)
torch.cuda.synchronize()
L = env.sim._L
L.drsim_debug_shard_times.restype = C.c_int
L.drsim_debug_shard_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
buf = np.zeros((1024, 16), dtype=np.uint64)
g = L.drsim_debug_shard_times(env.sim._h, buf.ctypes.data_as(C.c_void_p), 1024)
t = buf[:g].astype(np.int64)
t0 = t[:, 0].min()
names = ["start", "dep-wait done", "phase 1 done", "arrive+reduce done", "broadcast seen", "end"]
print(f"{g} CTAs; times in us relative to the first CTA's start (min / median / max over CTAs)")
for k, nm in enumerate(names):
    c = (t[:, k] - t0) / 1e3
    print(f"  {nm:20s} {c.min():8.2f} {np.median(c):8.2f} {c.max():8.2f}")
others = np.sort((np.delete(t[:, 3], int(np.argmax(t[:, 3]))) - t0) / 1e3)
print("partials published by the non-reducing CTAs (all warps done), us: p50 %.2f p90 %.2f p99 %.2f max %.2f" %
      (np.percentile(others, 50), np.percentile(others, 90), np.percentile(others, 99), others[-1]))
for k, nm in ((2, "phase-1 stamp"), (3, "partials out"), (5, "end")):
    order = np.argsort(-t[:, k])[:8]
    print(f"   latest {nm}: " + ", ".join(f"CTA {int(i)} {(t[i, k] - t0) / 1e3:.1f}" for i in order))
last = int(np.argmax(t[:, 3] - t[:, 2]))
print(f"   collection entered {(t[last, 11] - t0) / 1e3:.2f}, first pair of polls answered {(t[last, 13] - t0) / 1e3:.2f}, sweeps of thread 0: {t[last, 12]}")
for c in (0, 77, 155, 232, 148):
    print(f"   CTA {c}: phase-1 stamp {(t[c, 2] - t0) / 1e3:.2f} partials out {(t[c, 3] - t0) / 1e3:.2f} seen {(t[c, 4] - t0) / 1e3:.2f} end {(t[c, 5] - t0) / 1e3:.2f}")
print(f"   collection: thread 0 has its tiles {(t[last, 8] - t0) / 1e3:.2f}, all threads {(t[last, 9] - t0) / 1e3:.2f}, folded {(t[last, 10] - t0) / 1e3:.2f}")
print(f"reducing CTA {last}: phase 1 done {(t[last, 2] - t0) / 1e3:.2f}, partials collected + folded {(t[last, 6] - t0) / 1e3:.2f}, "
      f"published {(t[last, 7] - t0) / 1e3:.2f}, env planes stored {(t[last, 3] - t0) / 1e3:.2f}")
