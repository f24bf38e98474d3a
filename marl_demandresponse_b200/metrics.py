"""The reference's running metrics on the device, literally.

``ReferenceMetrics`` mirrors ``Metrics`` of ``server/app/services/metrics_service.py``: ``initialize`` :70-107,
``update`` :108-157, ``log`` :159-196 (the values, without wandb / the stray ``breakpoint()`` of :191), ``reset``
:222-235, ``update_rms`` :237-257 -- with the reference's own expressions, precedence slips included
(``temp_error = indoor_temp - target_temp / nb_agents``; the signal error divided by ``nb_agents**2`` once per
agent), so that what a rollout on the GPU logs is what the reference would have logged.  One launch of
``k_metrics_ref`` per step reads the state / reward planes of every replica and keeps the twelve cumulative fields
in a device tensor ``[R, 12]``; nothing is copied to the host until the values are asked for.

    m = ReferenceMetrics(env.sim, start_stats_from=0)
    for t in range(T):
        m.begin_step()            # obs_dict of the reference: signal / outdoor temperature / power BEFORE the step
        env.step(actions)
        m.end_step(t)             # next_obs_dict + rewards_dict: read from the planes
    m.log_values(time_steps_log)  # dict of [R] arrays, keys of metrics_service.py:176-187

The six running sums the step kernels keep themselves (``state["metrics"]``, reduced across ranks by
``distributed.reduce_rollout_metrics``) are the *intended* quantities (Ta - target, P - S per cluster) and cost
nothing extra; this class is the compatibility mode.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from . import _lib

FIELDS = ["cumul_avg_reward", "cumul_temp_offset", "cumul_temp_error", "max_temp_error", "cumul_signal_offset",
          "cumul_signal_error", "cumul_squared_error_temp", "cumul_OD_temp", "cumul_signal", "cumul_cons",
          "cumul_squared_error_sig", "cumul_squared_max_error_temp"]
_RESET = (0, 1, 2, 3, 4, 5, 7, 8, 9)   # Metrics.reset (:222-235) leaves the squared sums alone


class ReferenceMetrics:
    def __init__(self, sim, start_stats_from: int = 0, nb_time_steps: int = 0):
        import torch

        self.sim = getattr(sim, "sim", sim)   # a DrSim, or anything that carries one (BatchedEnv)
        self.start_stats_from = int(start_stats_from)
        self.nb_time_steps = int(nb_time_steps)
        self.nb_agents = self.sim.N
        dev = f"cuda:{self.sim.device}"
        self.acc = torch.zeros((self.sim.R, _lib.REF_METRIC_FIELDS), dtype=torch.float64, device=dev)
        self._prev = torch.zeros((self.sim.R, 3), dtype=torch.float64, device=dev)
        self._v = self.sim.views()

    def initialize(self, nb_agents: int = None, start_stats_from: int = None, nb_time_steps: int = None) -> None:
        """``Metrics.initialize`` (:70-107): every cumulative value back to zero."""
        if nb_agents is not None and int(nb_agents) != self.nb_agents:
            raise ValueError("nb_agents is the cluster size of the simulator")
        if start_stats_from is not None:
            self.start_stats_from = int(start_stats_from)
        if nb_time_steps is not None:
            self.nb_time_steps = int(nb_time_steps)
        self.acc.zero_()

    def begin_step(self) -> None:
        """Snapshot what the reference reads from ``obs_dict`` (:141-152): regulation signal, outdoor temperature and
        cluster power BEFORE the step."""
        self._prev[:, 0].copy_(self._v["signal"])
        self._prev[:, 1].copy_(self._v["od_temp"])
        self._prev[:, 2].copy_(self._v["power"])

    def end_step(self, time_step: int, stream=None) -> None:
        """``Metrics.update`` (:108-157) for every replica, from the planes the step just wrote."""
        sim = self.sim
        _lib.check(sim._L.drsim_metrics_update(sim._h, C.c_void_p(self._prev.data_ptr()), C.c_void_p(self.acc.data_ptr()),
                                               int(time_step >= self.start_stats_from), sim._stream(stream)))

    def reset(self) -> None:
        """``Metrics.reset`` (:222-235)."""
        self.acc[:, list(_RESET)] = 0.0

    def values(self) -> np.ndarray:
        """``[R, 12]`` host copy of the cumulative fields, order of ``FIELDS``."""
        return self.acc.cpu().numpy()

    def log_values(self, time_steps_log: int) -> Dict[str, np.ndarray]:
        """The dictionary ``Metrics.log`` builds (:176-187), one value per replica ("Mean signal error" repeats the
        offset there too)."""
        v = self.values() / float(time_steps_log)
        return {"Mean train return": v[:, 0], "Mean temperature offset": v[:, 1], "Mean temperature error": v[:, 2],
                "Mean signal error": v[:, 4], "Mean signal offset": v[:, 4], "Mean outside temperature": v[:, 7],
                "Mean signal": v[:, 8], "Mean consumption": v[:, 9]}

    def rms(self, time_step: int) -> Dict[str, np.ndarray]:
        """``Metrics.update_rms`` (:237-257)."""
        v = self.values()
        d = float(time_step - self.start_stats_from)
        return {"rmse_sig_per_ag": np.sqrt(v[:, 10] / d) / self.nb_agents,
                "rmse_temp": np.sqrt(v[:, 6] / (d * self.nb_agents)),
                "rms_max_error_temp": np.sqrt(v[:, 11] / d)}
