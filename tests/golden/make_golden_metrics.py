"""Generate ``tests/golden/metrics_*.json`` from the REAL reference ``Metrics`` class.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_golden_metrics.py

``server/app/services/metrics_service.py:108-157`` (``Metrics.update``) is fed, step by step, the observation
dicts of an existing golden trajectory (recorded from the running reference ``Environment`` by make_golden.py) the
way ``ControllerManager.start`` feeds it (controller_manager.py:171-176: ``update(obs_dict, next_obs_dict,
rewards_dict, t)``).  Its twelve cumulative fields are recorded after every step, ``update_rms`` at the end.  The
two modules the service imports that cannot load here are stubbed (``app.services.wandb_service``: wandb absent;
``app.utils.logger``: pydantic BaseSettings moved); neither takes part in the arithmetic.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from golden_util import GoldenCase  # noqa: E402
from oracle import refenv  # noqa: E402

FIELDS = ["cumul_avg_reward", "cumul_temp_offset", "cumul_temp_error", "max_temp_error", "cumul_signal_offset",
          "cumul_signal_error", "cumul_squared_error_temp", "cumul_OD_temp", "cumul_signal", "cumul_cons",
          "cumul_squared_error_sig", "cumul_squared_max_error_temp"]
CASES = [("c1_default_n10_bangbang", 3), ("random_n24_sinus_commonL2", 0), ("interp_n130_sampled", 10)]


def load_metrics():
    refenv.load()   # stubs + sys.path for app.core.environment
    logger_mod = types.ModuleType("app.utils.logger")
    logger_mod.logger = types.SimpleNamespace(info=lambda *a, **k: None, debug=lambda *a, **k: None)
    sys.modules["app.utils.logger"] = logger_mod
    wb = types.ModuleType("app.services.wandb_service")

    class WandbManager:   # noqa: D401 -- logging sink only
        def initialize(self): pass
        def log(self, *_a, **_k): pass

    wb.WandbManager = WandbManager
    sys.modules["app.services.wandb_service"] = wb
    import importlib.util

    if "app.services" not in sys.modules:
        pkg = types.ModuleType("app.services")
        pkg.__path__ = [os.path.join(refenv.REFERENCE_ROOT, "server/app/services")]
        sys.modules["app.services"] = pkg
    spec = importlib.util.spec_from_file_location(
        "app.services.metrics_service", os.path.join(refenv.REFERENCE_ROOT, "server/app/services/metrics_service.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.Metrics(WandbManager())


def obs_of(case, t):
    """The fields Metrics.update reads from the observation dicts BEFORE step t (t = 0: the reset observation)."""
    z, s0 = case.z, case.state0
    if t == 0:
        ta, sig, p, od = np.asarray(s0["t_air"])[0], float(np.asarray(s0["signal"])[0]), float(np.asarray(s0["power"])[0]), \
            float(np.asarray(s0["od_temp"])[0])
    else:
        ta, sig, p, od = z["t_air"][t - 1], float(z["signal"][t - 1]), float(z["power"][t - 1]), float(z["od_temp"][t - 1])
    tg = np.asarray(s0["target"])[0]
    return {k: {"indoor_temp": float(ta[k]), "target_temp": float(tg[k]), "reg_signal": sig, "cluster_hvac_power": p, "OD_temp": od}
            for k in range(case.N)}


def main():
    m = load_metrics()
    for name, start_from in CASES:
        case = GoldenCase(name)
        m.initialize(case.N, start_from, case.T)
        rows = []
        for t in range(case.T):
            m.update(obs_of(case, t), obs_of(case, t + 1), {k: float(case.z["rewards"][t][k]) for k in range(case.N)}, t)
            rows.append([float(getattr(m, f)) for f in FIELDS])
        m.update_rms(case.T)
        out = {"case": name, "start_stats_from": start_from, "fields": FIELDS, "per_step": rows,
               "rms": {"rmse_sig_per_ag": float(m.rmse_sig_per_ag), "rmse_temp": float(m.rmse_temp),
                       "rms_max_error_temp": float(m.rms_max_error_temp)}}
        path = os.path.join(HERE, f"metrics_{name}.json")
        json.dump(out, open(path, "w"))
        print(path, rows[-1][:4])


if __name__ == "__main__":
    main()
