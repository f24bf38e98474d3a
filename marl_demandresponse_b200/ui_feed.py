"""The UI feed of the reference, computed from the simulator's tensors (second half).

Mirror of ``ClientManagerService`` (``server/app/services/client_manager_service.py:28-287``): the
two payloads the reference emits once per step -- ``dataChange`` (twelve description strings,
:64-118) and ``houseChange`` (one small dict per house, :213-244) -- plus the graph series it
accumulates (:178-197).  Only the payloads: sockets, FastAPI and the Angular client stay where
they are; ``socket_manager`` is anything with ``async emit(endpoint, data)`` (the reference's own
``SocketManager`` fits).

``update_data`` accepts what the reference passes (the observation dicts) and, to skip building N
dicts per step, the environments themselves:

* the drop-in :class:`Environment` -- reads the host snapshot the step already pulled;
* a :class:`BatchedEnv` replica -- the aggregates come from one ``drsim_cluster_summary`` launch
  (fp64 sums in a fixed order), the per-house list from four row copies of that replica.
"""
from __future__ import annotations

import inspect
from typing import Any, Dict, List, Optional

import numpy as np

DESCRIPTION_KEYS = [          # client_manager_service.py:12-25
    "Number of HVAC",
    "Number of locked HVAC",
    "Outdoor temperature",
    "Average indoor temperature",
    "Average temperature difference",
    "Regulation signal",
    "Current consumption",
    "Consumption error (%)",
    "RMSE",
    "Mass temperature",
    "Target temperature",
    "Average temperature error",
]

_SERIES = ("temp_diff", "temp_err", "air_temp", "mass_temp", "target_temp", "outdoor_temp", "signal", "consumption")


class _Frame:
    """What one step contributes: cluster aggregates + the per-house columns of ``houseChange``."""

    __slots__ = ("n", "n_locked", "od", "mean_ta", "mean_diff", "mean_err", "mean_tm", "mean_target", "signal", "power",
                 "tm0", "target0", "on", "lockout", "sso", "ta", "target")


def _frame_from_columns(on, lockout, sso, ta, tm, target, od, signal, power) -> _Frame:
    f = _Frame()
    ta, tm, target = np.asarray(ta, np.float64), np.asarray(tm, np.float64), np.asarray(target, np.float64)
    f.n = int(ta.shape[0])
    f.n_locked = int(np.count_nonzero(lockout))
    f.od, f.signal, f.power = float(od), float(signal), float(power)
    d = ta - target
    f.mean_ta, f.mean_diff, f.mean_err = float(ta.mean()), float(d.mean()), float(np.abs(d).mean())
    f.mean_tm, f.mean_target = float(tm.mean()), float(target.mean())
    f.tm0, f.target0 = float(tm[0]), float(target[0])
    f.on, f.lockout, f.sso, f.ta, f.target = on, lockout, sso, ta, target
    return f


def _frame_from_obs(obs: Dict[int, Dict[str, Any]]) -> _Frame:
    ids = list(obs.keys())                       # pd.DataFrame(obs_dict).transpose() keeps dict order (:160)
    col = lambda k: [obs[i][k] for i in ids]     # noqa: E731
    o0 = obs[ids[0]]
    return _frame_from_columns([bool(x) for x in col("turned_on")], [bool(x) for x in col("lockout")], col("seconds_since_off"),
                               col("indoor_temp"), col("mass_temp"), col("target_temp"),
                               o0["OD_temp"], o0["reg_signal"], o0["cluster_hvac_power"])


def _frame_from_dropin(env) -> _Frame:
    s = env._snap
    return _frame_from_columns(s["on"][0].astype(bool).tolist(), s["lockout"][0].astype(bool).tolist(), s["sso"][0].tolist(),
                               s["t_air"][0], s["t_mass"][0], s["target"][0],
                               env.current_od_temp, s["signal"][0], s["power"][0])


def _frame_from_batched(env, replica: int, houses: bool) -> _Frame:
    import torch

    sim = env.sim
    summ = sim.cluster_summary()                 # [R, 8] on the device, one launch
    v = sim.views()
    r = int(replica)
    dev = v["temp_is_deviation"]
    ta_k, tm_k = ("dt_air", "dt_mass") if dev else ("t_air", "t_mass")
    head = torch.stack([summ[r, k] for k in range(8)] + [v["od_temp"][r], v["signal"][r], v["power"][r],
                                                        v[tm_k][r, 0].double(), v["target"][r, 0].double()]).cpu().numpy()
    f = _Frame()
    n = float(head[7])
    f.n, f.n_locked = int(n), int(head[0])
    f.mean_ta, f.mean_diff, f.mean_err = head[1] / n, head[2] / n, head[3] / n
    f.mean_tm, f.mean_target = head[4] / n, head[5] / n
    f.od, f.signal, f.power = float(head[8]), float(head[9]), float(head[10])
    f.target0 = float(head[12])
    f.tm0 = float(head[11]) + (f.target0 if dev else 0.0)
    if houses:
        flags = v["flags"][r].cpu().numpy()
        f.on, f.lockout = ((flags & 1) != 0).tolist(), ((flags & 2) != 0).tolist()
        f.sso = v["sso"][r].cpu().numpy().tolist()
        f.target = v["target"][r].double().cpu().numpy()
        f.ta = v[ta_k][r].double().cpu().numpy() + (f.target if dev else 0.0)
    else:
        f.on = f.lockout = f.sso = f.ta = f.target = None
    return f


class ClientFeed:
    """Same attributes and per-step behaviour as ``ClientManagerService``; see the module docstring."""

    def __init__(self, socket_manager: Any = None) -> None:
        self.socket_manager = socket_manager
        self.initialize_data(interface=socket_manager is not None)

    def initialize_data(self, interface: bool) -> None:                      # :125-145
        self.interface = interface
        self.description: Dict[int, Dict[str, str]] = {}
        for k in _SERIES:
            setattr(self, k, np.array([]))
        self.recent_signal = np.array([])
        self.recent_consumption = np.array([])
        self.houses_data: Dict[int, List[Dict[str, Any]]] = {}

    # ---- one step -------------------------------------------------------------------------
    def update(self, source: Any, time_step: int, replica: int = 0, houses: bool = True):
        """Synchronous core of ``update_data``: returns ``(houses_payload, description_payload)``."""
        if isinstance(source, dict):
            f = _frame_from_obs(source)
        elif hasattr(source, "_snap"):
            f = _frame_from_dropin(source)
        elif hasattr(source, "sim"):
            f = _frame_from_batched(source, replica, houses)
        else:
            raise TypeError("update_data takes the observation dicts, an Environment or a BatchedEnv")
        # graph series (:178-197)
        for k, x in (("temp_diff", f.mean_diff), ("temp_err", f.mean_err), ("air_temp", f.mean_ta), ("mass_temp", f.mean_tm),
                     ("target_temp", f.mean_target), ("outdoor_temp", f.od), ("signal", f.signal), ("consumption", f.power)):
            setattr(self, k, np.append(getattr(self, k), x))
        # description (:64-118, :199-211)
        with np.errstate(divide="ignore", invalid="ignore"):
            err_pct = (np.float64(f.signal) - np.float64(f.power)) / np.float64(f.signal) * 100
        values = [
            str(f.n),
            str(f.n_locked),
            str(round(f.od, 2)),
            str(round(f.mean_ta, 2)),
            str(round(f.mean_diff, 2)),
            str(f.signal),
            str(f.power),
            str(float(err_pct)),
            str(float(np.sqrt(np.mean((self.signal - self.consumption) ** 2)))),
            str(round(f.tm0, 2)),
            str(round(f.target0, 2)),
            str(float(np.mean(self.temp_err))),
        ]
        self.description[time_step] = dict(zip(DESCRIPTION_KEYS, values))
        # per-house list (:213-244)
        out: Optional[List[Dict[str, Any]]] = None
        if f.ta is not None:
            out = []
            ta, tg = f.ta.tolist(), f.target.tolist()
            for i in range(f.n):
                h: Dict[str, Any] = {"id": i}
                if f.on[i]:
                    h["hvacStatus"] = "ON"
                else:
                    h["hvacStatus"] = "Lockout" if f.lockout[i] else "OFF"
                    h["secondsSinceOff"] = f.sso[i]
                h["indoorTemp"] = ta[i]
                h["targetTemp"] = tg[i]
                h["tempDifference"] = ta[i] - tg[i]
                out.append(h)
            self.houses_data[time_step] = out
        return out, self.description[time_step]

    async def update_data(self, source: Any, time_step: int, replica: int = 0) -> None:   # :147-176
        houses, desc = self.update(source, time_step, replica)
        await self.log(emit=True, endpoint="houseChange", data=houses)
        await self.log(emit=True, endpoint="dataChange", data=desc)

    async def log(self, text: str = "", emit: bool = False, endpoint: str = "", data: Any = None) -> None:   # :246-270
        if self.interface and emit and endpoint != "" and self.socket_manager is not None:
            res = self.socket_manager.emit(endpoint, data)
            if inspect.isawaitable(res):
                await res

    async def get_state_at(self, time_step: int) -> None:                    # :272-287
        await self.log(emit=True, endpoint="timeStepData", data=self.description[time_step])
        await self.log(emit=True, endpoint="houseChange", data=self.houses_data[time_step])
