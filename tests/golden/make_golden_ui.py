"""Generate ``tests/golden/ui_feed_*.json`` from the REAL reference UI feed.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_golden_ui.py

The reference's ``ClientManagerService`` (``server/app/services/client_manager_service.py:28-287``)
is driven exactly as ``ControllerManager.start`` drives it (``controller_manager.py:129-189``:
``update_data(obs_dict, t)`` once per step) on a reference ``Environment`` reset from a Python
``random`` seed and stepped with the deadband bang-bang rule.  Its two socket payloads
(``dataChange`` = the description dict, ``houseChange`` = the per-house list) and its graph series
are recorded next to the observation dicts that produced them.  Two modules the service imports
cannot load here and are stubbed: ``app.utils.logger`` (pydantic ``BaseSettings`` moved) and
``app.services.socket_manager_service`` (python-socketio absent); neither takes part in the
arithmetic.
"""
from __future__ import annotations

import asyncio
import copy
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refenv  # noqa: E402
from oracle.np_oracle import deadband_bangbang  # noqa: E402

from make_golden import base_cfg  # noqa: E402

CASES = [
    dict(name="ui_feed_n10", seed=4, T=40, cfg=base_cfg(10)),
    dict(name="ui_feed_n33_sinus", seed=23, T=12, cfg=base_cfg(
        33, **{"power_grid_prop/signal_properties/mode": "sinusoidals",
               "cluster_prop/house_prop/hvac_prop/lockout_duration": 16})),
]


class _Socket:
    def __init__(self):
        self.sent = []

    async def emit(self, endpoint, data):
        self.sent.append((endpoint, copy.deepcopy(data)))


def load_service():
    logger_mod = types.ModuleType("app.utils.logger")
    logger_mod.logger = types.SimpleNamespace(info=lambda *a, **k: None, debug=lambda *a, **k: None)
    sys.modules["app.utils.logger"] = logger_mod
    sock_mod = types.ModuleType("app.services.socket_manager_service")
    sock_mod.SocketManager = _Socket
    sys.modules["app.services.socket_manager_service"] = sock_mod
    import importlib.util

    # load the one file (the package __init__ pulls in the whole server)
    if "app.services" not in sys.modules:
        pkg = types.ModuleType("app.services")
        pkg.__path__ = [os.path.join(refenv.REFERENCE_ROOT, "server/app/services")]
        sys.modules["app.services"] = pkg
    spec = importlib.util.spec_from_file_location(
        "app.services.client_manager_service",
        os.path.join(refenv.REFERENCE_ROOT, "server/app/services/client_manager_service.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, (np.bool_, bool)):
        return bool(x)
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    if hasattr(x, "isoformat"):
        return x.isoformat()
    return x


def run_case(ns, svc_mod, case):
    cfg = copy.deepcopy(case["cfg"])
    N = cfg["cluster_prop"]["nb_agents"]
    random.seed(case["seed"])
    env = ns.Environment(ns.EnvironmentProperties(**cfg))
    random.seed(case["seed"])
    obs = env.reset()
    sock = _Socket()
    svc = svc_mod.ClientManagerService(sock)
    svc.initialize_data(True)
    steps = []
    for t in range(case["T"]):
        a = deadband_bangbang(np.array([obs[i]["indoor_temp"] for i in range(N)]),
                              np.array([obs[i]["target_temp"] for i in range(N)]),
                              np.array([obs[i]["deadband"] for i in range(N)]),
                              np.array([obs[i]["turned_on"] for i in range(N)]))
        obs, _ = env.step({i: bool(a[i]) for i in range(N)})
        asyncio.run(svc.update_data(obs, t))
        keep = ("turned_on", "seconds_since_off", "lockout", "target_temp", "indoor_temp", "mass_temp",
                "cluster_hvac_power", "OD_temp", "reg_signal")
        steps.append(dict(
            actions=[int(x) for x in a],
            obs={i: {k: obs[i][k] for k in keep} for i in range(N)},
            description=svc.description[t],
            houses=svc.houses_data[t],
            emitted=[e for e, _ in sock.sent[-2:]],
        ))
    series = {k: getattr(svc, k).tolist() for k in
              ("temp_diff", "temp_err", "air_temp", "mass_temp", "target_temp", "outdoor_temp", "signal", "consumption")}
    return jsonable(dict(name=case["name"], seed=case["seed"], T=case["T"], env_prop=cfg,
                         description_keys=svc_mod.DESCRIPTION_KEYS, steps=steps, series=series))


def main():
    ns = refenv.load()
    svc_mod = load_service()
    for case in CASES:
        out = run_case(ns, svc_mod, case)
        path = os.path.join(HERE, case["name"] + ".json")
        with open(path, "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print(f"{case['name']:28s} {os.path.getsize(path) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
