"""Oracle restatement of the Monte-Carlo table generator (TEST INFRASTRUCTURE).

``server/v0/monteCarlo/monteCarlo.py:152-230`` driven through the fp64 NumPy oracle: one noise-free
single-house cluster per grid point, BangBangController, 75 steps, mean of the last 10 running
averages of the HVAC power.  The v0 and app environments share the thermal / HVAC model; the v0
specifics (HVAC initially off with ``seconds_since_off = lockout_duration``, constant outdoor
temperature, ``Ua`` multiplied rather than overwritten) are part of the injected initial state.

Pinned: ``tests/test_montecarlo_vs_reference.py`` evaluates the reference's own
``eval_parameters_bangbang_average_consumption`` (with the real legacy env and controller) on random grid
points and compares it with :func:`table_entries` (rtol 1e-9).
"""
from __future__ import annotations

import numpy as np

from .np_oracle import NpOracle, bangbang


def table_entries(prop: dict, state: dict, od_offset: np.ndarray, steps: int = 75, n_avg: int = 10) -> np.ndarray:
    R = np.asarray(state["epoch"]).shape[0]
    orc = NpOracle(prop, R)
    orc.set_state(state)
    total = np.zeros(R)
    avg = np.zeros(R)
    for i in range(steps):
        s = orc.state
        a = bangbang(s["t_air"], s["target"])          # v0/agents/bangbang_controllers.py:48-59
        orc.step(a, od_offset, None)
        total += orc.state["power"]
        if i >= steps - n_avg:
            avg += total / ((i + 1) * n_avg)           # monteCarlo.py:219-223
    return avg
