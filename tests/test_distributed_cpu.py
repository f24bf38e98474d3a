"""Host-side multi-rank logic on CPU (gloo, world_size 2): sharding arithmetic, placement-invariant
synthetic state, the end-of-rollout metric reduction and the rank-order layout of the gathered
partials that drsim_step_finish consumes."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_demandresponse_b200.batched import synthetic_state
from marl_demandresponse_b200.distributed import reduce_rollout_metrics, replica_shard
from marl_demandresponse_b200.sharded import halo_edge_houses, halo_lookup, house_shard


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


def test_halo_lookup_finds_every_ring_neighbour():
    """The halo layout of a house-sharded ring cluster (host restatement of k_reduce / k_obs): with every
    rank publishing the ids of its edge houses, each rank must resolve every ring neighbour of every house
    it owns either locally or to the right entry of an adjacent rank's block."""
    for n_global, world, c in ((1000, 2, 10), (1000, 8, 10), (96, 3, 5), (64, 4, 1), (20000, 8, 12)):
        blocks = [halo_edge_houses(n_global, r, world, c) for r in range(world)]
        L = c // 2
        for rank in range(world):
            lo, hi = house_shard(n_global, rank, world)
            assert hi - lo >= c
            for house in list(range(lo, min(hi, lo + c + 2))) + list(range(max(lo, hi - c - 2), hi)):
                for k in range(c):
                    want = (house - L + k) % n_global if k < L else (house + 1 + (k - L)) % n_global
                    where, got = halo_lookup(blocks, n_global, rank, world, c, house, k)
                    assert got == want, (n_global, world, c, rank, house, k, where)


def _halo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_global, c = 40, 6
        mine = torch.tensor(halo_edge_houses(n_global, rank, world, c), dtype=torch.float64)   # stands for halo_out
        gathered = torch.empty((world, c), dtype=torch.float64)
        dist.all_gather_into_tensor(gathered.view(-1), mine)
        lo, hi = house_shard(n_global, rank, world)
        bad = 0
        for house in range(lo, hi):
            for k in range(c):
                want = (house - c // 2 + k) % n_global if k < c // 2 else (house + 1 + (k - c // 2)) % n_global
                _, got = halo_lookup(gathered.tolist(), n_global, rank, world, c, house, k)
                bad += int(got != want)
        out.put((rank, bad))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_all_gather_layout():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get() for _ in range(2)) == [(0, 0), (1, 0)]
