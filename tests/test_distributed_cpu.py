"""Host-side multi-rank logic on CPU (gloo, world_size 2): sharding arithmetic, placement-invariant
synthetic state, the end-of-rollout metric reduction and the rank-order layout of the gathered
partials that drsim_step_finish consumes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_demandresponse_b200.batched import synthetic_state
from marl_demandresponse_b200.distributed import reduce_rollout_metrics, replica_shard
from marl_demandresponse_b200.sharded import house_shard


def test_replica_shard_partitions_exactly():
    for total in (1, 7, 16, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [replica_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (o0, n0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + n0 == o1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_house_shard_is_4_aligned_and_covers():
    for n in (1000, 1_000_000, 999_999, 37):
        for world in (1, 2, 4, 8):
            spans = [house_shard(n, r, world) for r in range(world)]
            covered = 0
            for lo, hi in spans:
                assert lo == covered and (lo % 4 == 0 or lo == hi)
                covered = hi
            assert covered == n


def test_synthetic_state_is_placement_invariant():
    prop = {"cluster_prop": {"nb_agents": 33}}
    whole = synthetic_state(prop, 10, seed=5)
    for rank, world in ((0, 2), (1, 2), (2, 3)):
        off, n = replica_shard(10, rank, world)
        part = synthetic_state(prop, n, seed=5, rep_offset=off)
        for k, v in part.items():
            assert np.array_equal(v, whole[k][off:off + n]), k


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # end-of-rollout metric reduction
        off, n = replica_shard(5, rank, world)
        g = torch.Generator().manual_seed(0)
        full = torch.rand((5, 6), generator=g, dtype=torch.float64)
        full[:, 0] = 10.0
        res = reduce_rollout_metrics(full[off:off + n])
        # gathered per-rank partials: [world, R, 6] in rank order, as drsim_step_finish expects
        acc = torch.full((3, 6), float(rank + 1), dtype=torch.float64)
        gathered = torch.empty((world * 3, 6), dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, acc)
        gathered = gathered.view(world, 3, 6)
        if rank == 0:
            out.put((res, gathered.numpy().copy(), full.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_metric_reduction_and_gather_layout():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res, gathered, full = out.get()
    assert res["replicas"] == 5 and res["replica_steps"] == 50
    np.testing.assert_allclose(res["mean_reward"], full[:, 1].sum() / 50)
    np.testing.assert_allclose(res["rms_signal_error"], (full[:, 5].sum() / 50) ** 0.5)
    assert np.all(gathered[0] == 1.0) and np.all(gathered[1] == 2.0)
