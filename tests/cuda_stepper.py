"""Adapter that drives the CUDA path (through the C ABI) with the stepper interface of
``golden_util.replay``."""
from __future__ import annotations

import numpy as np

from marl_demandresponse_b200 import DrSim, flatten_config


class CudaStepper:
    def __init__(self, env_prop, n_rep=1, precision="f32", path="auto", obs_layout="hand_engineered",
                 table=None, policy="external", noise="zero", seed=0, force_comm_table=None, host_api=True):
        mode = env_prop["cluster_prop"]["agents_comm_prop"]["mode"] if env_prop else "neighbours"
        use_table = (mode != "neighbours") if force_comm_table is None else force_comm_table
        cfg = flatten_config(env_prop, n_rep, precision, obs_layout, policy, noise, seed, path, comm_table=use_table)
        self.sim = DrSim(cfg)
        self.use_table = use_table and cfg.nb_comm > 0 and obs_layout == "hand_engineered"
        self.host_api = host_api
        self._table = None
        if table is not None:
            self.sim.set_interp_table(table)
        self._obs_valid = False

    def _set_table(self, table):
        if not self.use_table or table is None:
            return
        t = np.asarray(table, dtype=np.int32)
        if self._table is None or self._table.shape != t.shape or not np.array_equal(self._table, t):
            self.sim.set_comm_table(t)
            self._table = t.copy()
            self._obs_valid = False

    def set_state(self, st):
        self.sim.set_state(st)
        self._obs_valid = False

    reinject = set_state

    def step(self, actions, od_noise, perlin=None, interp_ids=None):
        import torch

        sim = self.sim
        if self.use_table and self._table is None:
            raise RuntimeError("neighbour table not set")
        if self.host_api:
            sim.step_host(np.asarray(actions, dtype=np.uint8).reshape(sim.R, sim.N), np.asarray(od_noise, dtype=np.float64),
                          None if perlin is None else np.asarray(perlin, dtype=np.float64),
                          None if interp_ids is None else np.asarray(interp_ids, dtype=np.int32))
        else:
            v = sim.views()
            v["actions"].copy_(torch.as_tensor(np.asarray(actions, dtype=np.uint8).reshape(sim.R, sim.N)))
            dev = v["actions"].device
            od = torch.as_tensor(np.asarray(od_noise, dtype=np.float64), device=dev)
            pe = None if perlin is None else torch.as_tensor(np.asarray(perlin, dtype=np.float64), device=dev)
            ids = None if interp_ids is None else torch.as_tensor(np.asarray(interp_ids, dtype=np.int32), device=dev).contiguous()
            sim.step(None, od, pe, ids)
            torch.cuda.synchronize()
        self._obs_valid = True
        return sim.views()["reward"].double().cpu().numpy()

    def get_state(self):
        return self.sim.get_state()

    def obs_vectors(self, table=None):
        import torch

        self._set_table(table)
        if not self._obs_valid:
            self.sim.refresh(False)
            self._obs_valid = True
        torch.cuda.synchronize()
        return self.sim.views()["obs"].double().cpu().numpy()
