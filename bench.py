#!/usr/bin/env python
"""bench.py -- house-steps/sec of the demand-response environment step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|c5]

Under ``torch.distributed.run`` one rank drives one GPU; replicas are sharded across ranks with
NO data-path collective (weak scaling: per-GPU work fixed).  Rank 0 prints ONE JSON line.

Workloads (BASELINE.json configs):
  c4  2,048 replicas per GPU x 1,000 houses, TarMAC observation layout (D = 10)      [default]
  c3  4,096 replicas (512 per GPU at N = 8) x 100 houses, hand-engineered obs (D = 50)
  c2  1 replica x 1,000 houses, interpolated base power, greedy-myopic on device (latency-bound)
The timed region of ``value`` has every input resident in HBM; ``e2e`` re-times the same workload
through the host-buffer C-ABI entry (``drsim_step_host``): pinned host actions -> device, per-replica
results -> host, every step.  ``--impl reference`` times the CPU port of the reference step
(``oracle/scalar_port.py``, the reference's own per-house Python loop restated) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "house-steps/sec"
UNIT = "house-steps/s"

WORKLOADS = {
    "c1": dict(name="C1: default cluster of 10 houses, 1 replica, DeadbandBangBang controller on the device, hand-engineered obs "
                    "(latency-bound: one tiny launch per step)",
               rep_per_gpu=1, n_houses=10, obs="hand_engineered", policy="deadband_bangbang"),
    "c2": dict(name="C2: 1,000-house cluster, 1 replica, interpolated base power (synthetic table, 100 sampled houses every 75 "
                    "steps), lock-out 40 s, greedy-myopic controller on the device (latency-bound)",
               rep_per_gpu=1, n_houses=1000, obs="hand_engineered", policy="greedy_myopic", base_mode="interpolation"),
    "c4": dict(name="C4: 2048 replicas/GPU x 1000 houses, TarMAC obs layout (D=10), external actions",
               rep_per_gpu=2048, n_houses=1000, obs="tarmac"),
    "c3": dict(name="C3: 4096 replicas/GPU x 100 houses, hand-engineered neighbour obs (D=50), external actions",
               rep_per_gpu=4096, n_houses=100, obs="hand_engineered"),
    "c5": dict(name="C5: ONE 1,000,000-house cluster split by houses across the GPUs, TarMAC obs layout, "
                    "per-step exchange of the per-rank aggregate-power partials over NVLink",
               rep_per_gpu=1, n_houses=1_000_000, obs="tarmac", sharded=True),
}


def env_prop_for(n_houses: int, base_mode: str = "constant") -> dict:
    prop = {
        "start_datetime": "2021-06-15T12:00:00", "start_datetime_mode": "fixed", "time_step": 4.0,
        "cluster_prop": {"nb_agents": n_houses, "house_prop": {"target_temp": 19.0}},
    }
    if base_mode != "constant":
        prop["power_grid_prop"] = {"base_power_props": {"mode": base_mode}}
    return prop


def algorithmic_bytes_per_house_step(real_bytes: int, obs_dim: int) -> int:
    """Every plane the step must touch once: state read+write
    (Ta, Tm, sso, flags), action, static parameters (6 thermal coefficients, target, capacity),
    reward, observation row."""
    state = 2 * real_bytes + 4 + 1
    static = 8 * real_bytes
    return state + 1 + static + state + real_bytes + obs_dim * real_bytes


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        rows = []
        for line in open(self.path).read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                rows.append(parts)
        os.unlink(self.path)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        out["samples"] = len(rows)
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
        try:
            out["sm_max_mhz"] = float(rows[0][2])
        except ValueError:
            pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------
# CPU arm: the scalar port of the reference step on the host cores
# ------------------------------------------------------------------------------------------
_WARM = {}


def _cpu_prop(n_houses: int, obs: str) -> dict:
    prop = env_prop_for(n_houses)
    prop["power_grid_prop"] = {"signal_properties": {"mode": "perlin"}}
    if obs == "tarmac":   # own-state features only: no neighbour messages are gathered
        prop["cluster_prop"]["agents_comm_prop"] = {"max_nb_agents_communication": 0}
    return prop


def _cpu_warm(n_houses: int, obs: str, seed: int):
    """Imports, environment construction and state injection of one worker -- everything that is NOT the step.
    Runs before any timed region (pool initializer / first call), once per process."""
    key = (n_houses, obs)
    if key not in _WARM:
        import numpy as np

        from marl_demandresponse_b200.batched import synthetic_state
        from oracle.scalar_port import ScalarEnv

        prop = _cpu_prop(n_houses, obs)
        env = ScalarEnv(prop)
        env.set_state(synthetic_state(prop, 1, seed=1234, rep_offset=seed, quirk_ua=False))
        rng = np.random.default_rng(seed)
        env.step(rng.random(n_houses) < 0.5, [0.0], [0.0])   # first-call costs (lazy imports, caches) stay outside
        _WARM[key] = (env, rng)
    return _WARM[key]


def _cpu_pool_init(n_houses: int, obs: str):
    _cpu_warm(n_houses, obs, os.getpid() & 0xFFFF)


def _cpu_worker(args):
    """``steps`` environment steps on this process' warm environment; returns the seconds spent stepping."""
    n_houses, obs, steps = args
    env, rng = _cpu_warm(n_houses, obs, os.getpid() & 0xFFFF)
    acts = rng.random((steps, n_houses)) < 0.5
    t0 = time.perf_counter()
    for t in range(steps):
        env.step(acts[t], [0.0], [0.0])
    return time.perf_counter() - t0


def cpu_numpy_rate(n_houses: int, obs: str, target_house_steps: float = 2e6) -> dict:
    """house-steps/s of the vectorised fp64 NumPy restatement (oracle/np_oracle.py) on one core: NOT how the
    reference computes (it loops over Python objects), reported next to the port so the CPU side is not
    only a slow baseline."""
    import numpy as np

    from marl_demandresponse_b200.batched import synthetic_state
    from oracle.np_oracle import NpOracle, from_epoch

    prop = _cpu_prop(n_houses, obs)
    R = max(1, min(64, int(65536 // n_houses)))
    steps = max(2, int(target_house_steps / (R * n_houses)))
    o = NpOracle(prop, R)
    o.set_state(synthetic_state(prop, R, seed=1234, quirk_ua=False))
    o.power_grid_step([from_epoch(e) for e in o.state["epoch"]], np.zeros(R))
    acts = np.random.default_rng(0).random((steps, R, n_houses)) < 0.5
    zero = np.zeros(R)
    t0 = time.perf_counter()
    for t in range(steps):
        o.step(acts[t], zero, zero)
        o.obs_vectors()
    dt = time.perf_counter() - t0
    return {"value": R * n_houses * steps / dt, "unit": UNIT, "cores": 1, "kind": "vectorised NumPy restatement (not the reference's implementation)",
            "sample": f"{R} replica(s) x {n_houses} houses x {steps} steps of oracle/np_oracle.py incl. observation vectors"}


def cpu_model() -> str:
    """CPU model of the box the baseline ran on (SURVEY 8d asks for it next to the core count)."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class CpuPort:
    """``procs`` warm worker processes, each owning one single-cluster replica of the scalar port.  A timed
    call covers stepping only: the workers import, build and warm their environment in the pool initializer,
    and the rate is computed from the slowest worker's own stepping time (procs * N * steps / max dt) -- pool
    dispatch and pickling are not the reference's step either."""

    def __init__(self, n_houses: int, obs: str, procs: int):
        self.n, self.obs, self.procs = n_houses, obs, procs
        self.pool = None
        if procs > 1:
            import multiprocessing as mp

            self.pool = mp.get_context("spawn").Pool(procs, initializer=_cpu_pool_init, initargs=(n_houses, obs))
            self.pool.map(_cpu_worker, [(n_houses, obs, 1)] * procs)   # every worker is up and warm
        else:
            _cpu_warm(n_houses, obs, os.getpid() & 0xFFFF)

    def rate(self, steps: int) -> dict:
        if self.pool is None:
            dts = [_cpu_worker((self.n, self.obs, steps))]
        else:
            dts = self.pool.map(_cpu_worker, [(self.n, self.obs, steps)] * self.procs, chunksize=1)
        rate = self.procs * self.n * steps / max(dts)
        return {"value": rate, "unit": UNIT, "cores": self.procs, "kind": "port", "cpu_model": cpu_model(),
                "host_cores": os.cpu_count(), "seconds": max(dts),
                "sample": f"{self.procs} replica(s) x {self.n} houses x {steps} steps of oracle/scalar_port.py "
                          f"(per-house Python loop restating the reference step), warm workers, stepping time only"}

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


def cpu_port_rate(n_houses: int, obs: str, steps: int, procs: int) -> dict:
    """house-steps/s of the scalar port: ``procs`` independent single-cluster replicas."""
    port = CpuPort(n_houses, obs, procs)
    try:
        return port.rate(steps)
    finally:
        port.close()


def workload_config(wl: dict, world: int, obs_dim: int, exchange: str = "peer", flush_l2: bool = False) -> dict:
    """The ``config`` object of the JSON line -- the SAME for both arms (it names the workload, not the machinery)."""
    R, N = wl["rep_per_gpu"], wl["n_houses"]
    sharded = bool(wl.get("sharded"))
    bytes_hs = algorithmic_bytes_per_house_step(4, obs_dim)
    n_local = -(-N // world) if sharded else N
    ws = R * n_local * bytes_hs
    on_device_policy = wl.get("policy", "external") != "external"
    return {"workload": wl["name"], "replicas_per_gpu": R, "houses_per_cluster": N, "obs_dim": obs_dim,
            "parallelism": (f"house-sharded x{world}, per-step exchange of 48 B/rank via {exchange if world > 1 else 'none'}" if sharded
                            else f"replica-sharded x{world}, no per-step collective"),
            "l2": f"working set {ws / 1e6:.0f} MB per step per GPU vs 126 MB L2"
                  + ("" if ws > 126e6 else "; L2 flushed between timed steps by a 256 MB write (flush time excluded: one "
                     "event pair per step)" if flush_l2 else "; fits L2 (latency-bound workload, see DESIGN.md)"),
            "actions": (f"on-device controller ({wl['policy']})" if on_device_policy else
                        "4 rotating fixed-seed Bernoulli(0.5) u8 tensors (policy cost excluded)")}


def obs_dim_of(wl: dict) -> int:
    """Observation row width of a workload (utils/norm.py:178-218: 10 own-state features + 4 per neighbour)."""
    if wl["obs"] == "tarmac":
        return 10
    return 10 + 4 * min(10, wl["n_houses"] - 1)


def run_reference_arm(args, wl) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    n = wl["n_houses"] if not wl.get("sharded") else 1000   # C5 per-house cost is extrapolated from a 1000-house cluster
    per_step = max(300, int(round(3e5 / n)))   # >= 300 environment steps per worker and timed sample (~3 s of stepping)
    port = CpuPort(n, wl["obs"], procs)        # workers import / build / warm up here, outside every timed sample
    one = CpuPort(n, wl["obs"], 1).rate(per_step)   # the same sample on one core: the all-core rate should be ~cores x this
    t0 = time.perf_counter()
    rates = []
    for i in range(args.warmup + args.steps):
        r = port.rate(per_step)
        if i >= args.warmup:
            rates.append(r)
        if time.perf_counter() - t0 > 200:
            break
    port.close()
    rates = rates or [r]
    # steps-weighted mean: total house-steps over total (slowest-worker) stepping time
    value = procs * n * per_step * len(rates) / sum(x["seconds"] for x in rates)
    cb = dict(rates[-1])
    cb["value"] = value
    cb["one_core"] = one["value"]
    cb["parallel_efficiency"] = value / (procs * one["value"])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
        "warmup": args.warmup, "ms_per_step": 1e3 * procs * n * per_step / value, "higher_is_better": True,
        "scaling": "strong" if wl.get("sharded") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": workload_config(wl, max(world, args.gpus), obs_dim_of(wl), args.exchange, args.flush_l2),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def source_hash() -> str:
    """sha256 over the CUDA sources + the public header: what a committed ncu capture must match to be current."""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "marl_demandresponse_b200", "csrc")
    for f in sorted(os.listdir(csrc)) + [os.path.join("..", "..", "include", "drsim.h")]:
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def committed_traffic(workload: str) -> dict:
    """DRAM bytes per launch of the workload's dominant kernel from the committed ncu capture
    (profiles/r2_traffic.json, written by profiles/tools/ncu_traffic.py); `traffic_stale` says whether the kernels
    have changed since it was taken."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        t = d.get(workload)
        if not t:
            return {"traffic": None}
        k = float(t.get("steps_per_launch", 1))   # a launch of the in-kernel step loop advances 64 steps: traffic is per STEP
        return {"traffic": (t["dram_bytes_read"] + t["dram_bytes_write"]) / k, "traffic_read": t["dram_bytes_read"] / k,
                "traffic_write": t["dram_bytes_write"] / k, "traffic_launches_averaged": t.get("launches"),
                "traffic_per": "step" + (f" (launch total / {int(k)} steps)" if k > 1 else " (= launch)"),
                "traffic_stale": d.get("source_hash") != source_hash()}
    except Exception:  # noqa: BLE001
        return {"traffic": None}


class GpuWorkload:
    """One workload on this rank's GPU: environment, rotating action tape, timing helpers."""

    def __init__(self, name: str, args, rank: int, world: int, local: int):
        import numpy as np
        import torch

        from marl_demandresponse_b200 import BatchedEnv

        self.wl = wl = WORKLOADS[name]
        self.name, self.args, self.rank, self.world, self.local = name, args, rank, world, local
        self.dev = dev = torch.device("cuda", local)
        R, N = wl["rep_per_gpu"], wl["n_houses"]
        self.R, self.N = R, N
        self.sharded = bool(wl.get("sharded"))
        if self.sharded:
            from marl_demandresponse_b200.sharded import ShardedClusterEnv

            self.env = ShardedClusterEnv(env_prop_for(N), R, rank=rank, world=world, device=local, precision="f32",
                                         obs_layout=wl["obs"], noise="philox", seed=1234, exchange=args.exchange)
            self.env.reset()
            self.n_local = self.env.hi - self.env.lo
        else:
            table = None
            if wl.get("base_mode") == "interpolation":
                table = interp_table_for_bench()
            self.env = BatchedEnv(env_prop_for(N, wl.get("base_mode", "constant")), R, device=local, precision="f32",
                                  obs_layout=wl["obs"], policy=wl.get("policy", "external"), noise="philox", seed=1234,
                                  rep_offset=rank * R, interp_table=table)
            self.env.reset()
            self.n_local = N
        self.D = self.env.sim.D
        self.on_device_policy = wl.get("policy", "external") != "external"
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        self.n_act = 4
        Ns = self.env.sim.Ns
        # rotating action tape [4][R][Ns] (house axis padded to the plane stride), fixed-seed Bernoulli(0.5)
        self.tape = torch.zeros((self.n_act, R, Ns), dtype=torch.uint8, device=dev)
        self.tape[:, :, :self.n_local] = (torch.rand((self.n_act, R, self.n_local), device=dev, generator=g) < 0.5).to(torch.uint8)
        self.acts = [self.tape[i, :, :self.n_local] for i in range(self.n_act)]
        self.total_houses = R * N if self.sharded else world * R * N
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush_l2 else None

    # -- timing ------------------------------------------------------------------------------
    def barrier(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, one_call=False):
        """CUDA events around `steps` steps (barrier + synchronize on both sides), MAX over ranks, in ms.  With
        --flush-l2 every step is bracketed by its own event pair and a 256 MB write runs between the pairs (its
        time is not counted)."""
        import torch
        import torch.distributed as dist

        self.barrier()
        if self.flush_buf is not None and not one_call:
            pairs = []
            for i in range(steps):
                self.flush_buf.fill_(i & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(i)
                e1.record()
                pairs.append((e0, e1))
            self.barrier()
            t = sum(a.elapsed_time(b) for a, b in pairs)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if one_call:
                fn(steps)      # the whole run in one C call (drsim_run_tape)
            else:
                for i in range(steps):
                    fn(i)
            e1.record()
            self.barrier()
            t = e0.elapsed_time(e1)
        ms = torch.tensor([t], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_dev(self, i):
        self.env.step(None if self.on_device_policy else self.acts[i % self.n_act])

    def run_dev(self, k):
        """K device-resident steps enqueued by ONE C call: on-device controller, or the rotating action tape
        (step k replays plane k % 4) -- no host-language round trip, no launch jitter inside the timed region."""
        if self.on_device_policy:
            self.env.run(k)
        else:
            self.env.run(k, self.tape, rotate=True)

    def device_resident(self, steps, warmup):
        """(ms per step, launches inside the timed region)"""
        for i in range(max(3, warmup)):
            self.step_dev(i)
        one_call = self.flush_buf is None and (self.sharded is False or self.args.exchange == "peer" or self.world == 1)
        if one_call:
            self.run_dev(8)
        l0 = self.env.sim.launch_count
        ms = self.timed(self.run_dev, steps, one_call=True) if one_call else self.timed(self.step_dev, steps)
        return ms / steps, self.env.sim.launch_count - l0, one_call


def interp_table_for_bench():
    """The interpolation table of BASELINE config 2: GENERATED by the Monte-Carlo table generator (SURVEY 8f-1,
    v0/monteCarlo/monteCarlo.py:152-230 on the GPU) on a coarse sub-grid when the GPU is free for it, else the seeded
    stand-in for the reference's missing mergedGridSearchResultFinal.npy."""
    import numpy as np

    try:
        from marl_demandresponse_b200.montecarlo import generate_table

        return generate_table()
    except Exception:  # noqa: BLE001
        return np.random.default_rng(2024).uniform(0.0, 6000.0, 3 * 3 * 3 * 3 * 9 * 5 * 8 * 2 * 12 * 6)


def run_gpu_arm(args, wl) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    w = GpuWorkload(args.workload, args, rank, world, local)
    env, R, N, D = w.env, w.R, w.N, w.D
    sharded, on_device_policy = w.sharded, w.on_device_policy
    n_local = w.n_local

    # ---- device-resident: `value` -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # nvidia-smi needs ~0.1 s to deliver its first sample: start it before the warm-up
    ms_step, launches, one_call = w.device_resident(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else {}
    # the same drsim_run_tape call with the in-kernel step loop switched off: one launch per step (the definition of
    # `value` before the loop existed), reported next to it under roofline.per_step_launch
    streamed = one_call and launches * 2 < args.steps
    ms_launch = launches_launch = None
    if streamed:
        os.environ["DRSIM_NO_STREAM"] = "1"
        try:
            ms_launch, launches_launch, _ = w.device_resident(min(args.steps, 2000), args.warmup)
        finally:
            del os.environ["DRSIM_NO_STREAM"]

    # ---- end to end through the host-buffer C-ABI entry: `e2e` (full result) and `e2e_summary` -------------
    acts_host = [a.contiguous().cpu().pin_memory() for a in w.acts]
    env_out = torch.zeros((R, 4), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 300))
    e2e = None
    if sharded:
        def step_host(i):  # host actions in (pinned), per-replica results out, around the sharded step
            env.state["actions"].copy_(acts_host[i % w.n_act], non_blocking=True)
            env.step(None)
            env_out[:, 0].copy_(env.state["power"], non_blocking=True)
            torch.cuda.synchronize()
        step_full = None
    elif on_device_policy:
        step_host = lambda i: env.step_host(None, env_out)   # no actions to send: per-replica results out, sync
        step_full = None
    else:
        step_host = lambda i: env.step_host(acts_host[i % w.n_act], env_out)
        rew_host = torch.zeros((R, N), dtype=torch.float32).pin_memory()
        obs_host = torch.zeros((R, N, D), dtype=torch.float32).pin_memory() if D else None
        step_full = lambda i: env.step_host(acts_host[i % w.n_act], env_out, reward_out=rew_host, obs_out=obs_host)
    for i in range(max(50, args.warmup)):   # in-place reads of a pinned buffer run slower for the first few hundred steps (host-page mappings warm up)
        step_host(i)
    ms_sum = w.timed(step_host, e2e_steps) / e2e_steps
    h2d = 0 if on_device_policy else w.total_houses
    summary = {"value": w.total_houses / (ms_sum * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": world * R * 4 * 8, "ms_per_step": ms_sum,
               "returns": "per-replica [R][4] doubles: cluster power, signal, outdoor temperature, mean reward"}
    if step_full is not None:
        for i in range(5):
            step_full(i)
        ms_full = w.timed(step_full, e2e_steps) / e2e_steps
        e2e = {"value": w.total_houses / (ms_full * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": world * (R * 4 * 8 + R * N * 4 + R * N * D * 4), "ms_per_step": ms_full,
               "returns": "what Environment.step returns (environment.py:108): per-agent observation rows [R][N][D] f32 + "
                          "rewards [R][N] f32 into pinned host buffers, plus the [R][4] summary",
               "api": "BatchedEnv.step_host(actions, env_out, reward_out, obs_out) -> drsim_step_host_full: pinned host "
                      "actions in, full result out, one stream sync per step"}
    else:
        e2e = dict(summary)
        e2e["note"] = ("on-device controller: no actions to send, the [R][4] summary is the step's result" if on_device_policy else
                       "house-sharded cluster: actions H2D + step + cluster power D2H + sync through the torch views")

    # ---- the whole rollout transition on the device (nothing crosses PCIe) ---------------------
    rollout_line = None
    if not args.no_rollout and not sharded and not on_device_policy and 1 <= D <= 64:
        torch.manual_seed(4 + rank)
        fc = torch.nn.ModuleList([torch.nn.Linear(D, 100), torch.nn.Linear(100, 100), torch.nn.Linear(100, 2)]).to(dev)
        weights = type(env).actor_weights(fc)
        from marl_demandresponse_b200.rollout import RolloutBuffer

        seg = 8   # transitions per collected segment
        buf = RolloutBuffer(env, seg)
        r_steps = max(seg, min(args.steps, 160) // seg * seg)
        flops = 2.0 * R * N * (D * 100 + 100 * 100 + 100 * 2)
        per = {}
        for prec in ("tf32x3", "tf32"):   # fp32-grade probabilities (three passes per product) / one TF32 pass
            env.collect(weights, buf, precision=prec)
            ms_pol = w.timed(lambda i: env.policy_step(weights, precision=prec), r_steps)
            ms_roll = w.timed(lambda i: env.collect(weights, buf, precision=prec), r_steps // seg)
            per[prec] = {"us_per_transition": 1e3 * ms_roll / r_steps, "actor_us": 1e3 * ms_pol / r_steps,
                         "actor_tflops": (3 if prec == "tf32x3" else 1) * flops * r_steps / (ms_pol * 1e-3) / 1e12,
                         "value": world * R * N * r_steps / (ms_roll * 1e-3)}
        rollout_line = {"value": per["tf32x3"]["value"], "unit": "agent-steps/s",
                        "us_per_transition": per["tf32x3"]["us_per_transition"], "actor_us": per["tf32x3"]["actor_us"],
                        "actor_tflops": per["tf32x3"]["actor_tflops"],
                        "actor_precision": "tf32x3: every operand split into hi + lo TF32 halves, three tcgen05 passes per "
                                           "product, probabilities within ~1e-6 of an fp32 forward (the parity-grade default)",
                        "single_pass_tf32": per["tf32"],
                        "what": "device-resident rollout: MAPPO.select_actions (actor + categorical draw, mappo.py:83-97), "
                                "Environment.step and MAPPO.store_transition (mappo.py:105-127) per transition in one C call "
                                "(drsim_rollout_transition); state / action / probability / reward / next state land in a "
                                "device buffer written by the kernels themselves, nothing crosses PCIe",
                        "transition_bytes_stored_per_step": world * R * env.sim.Ns * (D * 4 + 1 + 4 + 4),
                        "actor": f"tcgen05 TF32, [{D} -> 100 -> 100 -> 2], random-init weights", "steps": r_steps}
        del buf

    # ---- BASELINE config 5 next to the main workload: ONE 1M-house cluster split by houses over the ranks ----
    sharded_line = None
    if not sharded and not args.no_c5:
        try:
            del w.tape, w.acts
            c5 = GpuWorkload("c5", args, rank, world, local)
            c5_steps = max(3, min(args.steps, 2000))
            ms5, l5, oc5 = c5.device_resident(c5_steps, args.warmup)
            b5 = algorithmic_bytes_per_house_step(4, c5.D)
            c5.env.sim.peer_status()
            sharded_line = {"workload": c5.wl["name"], "value": c5.total_houses / (ms5 * 1e-3), "unit": UNIT, "us_per_step": ms5 * 1e3,
                            "houses_per_gpu": c5.n_local, "scaling": "strong", "exchange": args.exchange if world > 1 else "none",
                            "kernel": "k_shard (one persistent kernel per step: update, reduction, NVLink push of the "
                                      "partial sums, wait, epilogue, rewards, rows)",
                            "launches_per_step": l5 / c5_steps, "steps": c5_steps,
                            "roofline_frac_per_gpu": c5.R * c5.n_local * b5 / (ms5 * 1e-3) / 1e9 / peak_gbs()}
            del c5
        except Exception as e:  # noqa: BLE001 -- an extra, never a reason to lose the line
            sharded_line = {"error": str(e)[:300]}

    value = w.total_houses / (ms_step * 1e-3)
    # end-of-rollout metric reduction (one all-reduce of a handful of fp64 sums)
    from marl_demandresponse_b200.distributed import reduce_rollout_metrics

    rollout = reduce_rollout_metrics(env.state["metrics"])

    if rank == 0:
        peak = peak_gbs()
        bytes_hs = algorithmic_bytes_per_house_step(4, D)
        achieved = R * n_local * bytes_hs / (ms_step * 1e-3) / 1e9
        variant = env.sim.fused_info()["variant"]
        kernel = ("k_shard" if (sharded or variant == "none") else
                  {"staged": "k_fused_tma", "staged_rows": "k_fused_rows", "direct": "k_fused_direct", "chunked": "k_fused"}[variant])
        cpu = cpu_port_rate(N, wl["obs"], max(1, int(1e6 / N)), 1) if world == 1 and not args.no_cpu else None   # ~10 s of CPU work
        cfg = workload_config(wl, world, D, args.exchange, args.flush_l2)
        cfg["timed_region"] = (f"{args.steps} steps enqueued by one drsim_run_tape call (rotating 4-plane action tape)" if one_call
                               else f"{args.steps} Python -> C step calls")
        if streamed:
            cfg["timed_region"] += ("; the step loop runs inside the kernel (one launch per 64-step block of schedule records: a "
                                    "CTA's inputs of step k+1 come only from itself, so the steps of a block have no boundary); "
                                    "roofline.per_step_launch = the same call with one launch per step")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "rollout_metrics": rollout,
            "e2e": e2e, "e2e_summary": summary,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kernel,
                         "achieved": achieved, "peak": peak,
                         **({"note": "latency-bound workload (one small cluster): the fraction is reported for completeness"}
                            if R * n_local < 100000 else {}),
                         "unit": "GB/s", "frac": achieved / peak, **committed_traffic(args.workload),
                         "algorithmic_bytes_per_step": R * n_local * bytes_hs,
                         "steps_per_launch": min(64, args.steps) if streamed else 1,
                         "algorithmic_bytes_per_launch": R * n_local * bytes_hs * (min(64, args.steps) if streamed else 1),
                         "bytes_per_house_step": bytes_hs,
                         **({"per_step_launch": {"ms_per_step": ms_launch, "value": w.total_houses / (ms_launch * 1e-3),
                                                 "achieved": R * n_local * bytes_hs / (ms_launch * 1e-3) / 1e9,
                                                 "frac": R * n_local * bytes_hs / (ms_launch * 1e-3) / 1e9 / peak,
                                                 "gpu_launches": launches_launch, "steps": min(args.steps, 2000),
                                                 **committed_traffic(args.workload + "_per_step_launch")}}
                            if streamed else {}),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else "fallback 6650"},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
            try:
                line["cpu_baseline"]["vectorised"] = cpu_numpy_rate(N if not sharded else min(N, 65536), wl["obs"])
            except Exception as e:  # noqa: BLE001 -- an extra, never a reason to lose the line
                line["cpu_baseline"]["vectorised"] = {"error": str(e)[:200]}
        if rollout_line:
            line["rollout"] = rollout_line
        if sharded_line:
            line["sharded_cluster"] = sharded_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:  # noqa: BLE001
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--flush-l2", action="store_true", help="write a 256 MB buffer between timed steps")
    ap.add_argument("--no-rollout", action="store_true",
                    help="skip the device-resident rollout transition (MA-PPO actor + draw + env step), the \"rollout\" key")
    ap.add_argument("--rollout", action="store_true", help="(accepted for compatibility: the rollout line is on by default)")
    ap.add_argument("--no-c5", action="store_true",
                    help="skip the house-sharded 1M-house cluster (BASELINE config 5) timed next to the main workload, the "
                         "\"sharded_cluster\" key")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="c5 only: per-step exchange of the aggregate-power partials (peer-memory stores vs NCCL all-gather)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
