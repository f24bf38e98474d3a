"""Loader + instrumentation for the REAL reference environment (authoring container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box: everything here is
guarded by :func:`available`.  It is used (a) by ``tests/golden/make_golden.py`` to generate
the committed golden trajectories and (b) by ``tests/test_oracle_vs_reference.py`` to pin the
NumPy / scalar restatements against the running reference.

The reference env needs two third-party modules that are absent here
(``server/app/core/environment/power_grid/perlin.py:1-2``): ``matplotlib`` and
``perlin_noise``.  Both are stubbed in ``sys.modules``; the stub ``PerlinNoise.noise`` is a
deterministic smooth function so that the *consumer* arithmetic (``perlin.py:41-56``,
``signal_calculator.py:100-115``) is exercised, while the values themselves stay
"parity unpinned" (SURVEY.md section 8c).
"""
from __future__ import annotations

import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DRSIM_REFERENCE_ROOT", "/root/reference")
_SERVER = os.path.join(REFERENCE_ROOT, "server")


def available() -> bool:
    return os.path.isdir(os.path.join(_SERVER, "app", "core", "environment"))


class StubPerlinNoise:
    """Stand-in for ``perlin_noise.PerlinNoise`` (values arbitrary but deterministic)."""

    def __init__(self, octaves=1, seed=1):
        self.octaves = float(octaves)
        self.seed = float(seed)

    def noise(self, x):
        return 0.5 * math.sin(self.octaves * float(x) * 1.7 + 6.0 * self.seed)


def _install_stubs() -> None:
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "perlin_noise" not in sys.modules:
        pn = types.ModuleType("perlin_noise")
        pn.PerlinNoise = StubPerlinNoise
        sys.modules["perlin_noise"] = pn


def load():
    """Import the reference env modules; returns a namespace of the pieces the tests use."""
    if not available():
        raise RuntimeError(f"reference not present under {REFERENCE_ROOT}")
    _install_stubs()
    if _SERVER not in sys.path:
        sys.path.insert(0, _SERVER)
    from app.core.environment import environment as env_mod
    from app.core.environment.cluster import cluster as cluster_mod
    from app.core.environment.cluster import agent_communication_builder as comm_mod
    from app.core.environment.environment_properties import EnvironmentProperties
    from app.core.environment.power_grid import interpolation as interp_mod
    from app.core.environment.power_grid import perlin as perlin_mod
    from app.core.environment.power_grid import signal_calculator as sig_mod
    from app.utils import norm as norm_mod

    ns = types.SimpleNamespace(
        env_mod=env_mod,
        cluster_mod=cluster_mod,
        comm_mod=comm_mod,
        interp_mod=interp_mod,
        perlin_mod=perlin_mod,
        sig_mod=sig_mod,
        norm_mod=norm_mod,
        Environment=env_mod.Environment,
        EnvironmentProperties=EnvironmentProperties,
        norm_state_dict=norm_mod.norm_state_dict,
    )
    return ns


class Recorder:
    """Wraps the per-step noise sources of the reference so their draws can be logged.

    * ``random.gauss`` as seen by ``environment.py:158``  -> ``od_noise``
    * ``Perlin.calculate_noise`` (``perlin.py:41``)        -> ``perlin``
    * ``random.choices`` as seen by ``interpolation.py:223`` -> ``interp_ids``
    * ``AgentCommunicationBuilder.get_random_sample`` (``agent_communication_builder.py:191``)
      -> ``comm_samples`` (one list per call, in call order)
    """

    def __init__(self, ns):
        import random as _random

        self.ns = ns
        self.od_noise = []
        self.perlin = []
        self.interp_ids = []
        self.comm_samples = []
        rec = self

        class _EnvRandom:
            def __getattr__(self, name):
                return getattr(_random, name)

            @staticmethod
            def gauss(mu, sigma):
                v = _random.gauss(mu, sigma)
                rec.od_noise.append(v)
                return v

        class _InterpRandom:
            def __getattr__(self, name):
                return getattr(_random, name)

            @staticmethod
            def choices(pop, k=1):
                v = _random.choices(pop, k=k)
                rec.interp_ids.append(list(v))
                return v

        self._saved = (
            ns.env_mod.random,
            ns.interp_mod.random,
            ns.perlin_mod.Perlin.calculate_noise,
            ns.comm_mod.AgentCommunicationBuilder.get_random_sample,
        )
        ns.env_mod.random = _EnvRandom()
        ns.interp_mod.random = _InterpRandom()
        orig_calc = self._saved[2]
        orig_sample = self._saved[3]

        def calc(self_, x):
            v = orig_calc(self_, x)
            rec.perlin.append(v)
            return v

        def sample(self_, agent_id):
            v = orig_sample(self_, agent_id)
            rec.comm_samples.append(list(v))
            return v

        ns.perlin_mod.Perlin.calculate_noise = calc
        ns.comm_mod.AgentCommunicationBuilder.get_random_sample = sample

    def close(self):
        ns = self.ns
        ns.env_mod.random = self._saved[0]
        ns.interp_mod.random = self._saved[1]
        ns.perlin_mod.Perlin.calculate_noise = self._saved[2]
        ns.comm_mod.AgentCommunicationBuilder.get_random_sample = self._saved[3]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def extract_state(env) -> dict:
    """Flatten a reference ``Environment`` into the oracle's state dict (R = 1)."""
    import numpy as np

    from .np_oracle import to_epoch

    b = env.cluster.buildings
    arr = lambda f, dt=np.float64: np.array([[f(x) for x in b]], dtype=dt)
    st = {
        "t_air": arr(lambda x: x.indoor_temp),
        "t_mass": arr(lambda x: x.current_mass_temp),
        "target": arr(lambda x: x.init_props.target_temp),
        "Ua": arr(lambda x: x.init_props.Ua),
        "Ca": arr(lambda x: x.init_props.Ca),
        "Cm": arr(lambda x: x.init_props.Cm),
        "Hm": arr(lambda x: x.init_props.Hm),
        "cap": arr(lambda x: x.hvac.init_props.cooling_capacity),
        "on": arr(lambda x: bool(x.hvac.turned_on), bool),
        "lockout": arr(lambda x: bool(x.hvac.lockout), bool),
        "sso": arr(lambda x: x.hvac.seconds_since_off, np.int64),
        "epoch": np.array([to_epoch(env.date_time)], dtype=np.int64),
        "od_temp": np.array([float(env.current_od_temp)]),
        "signal": np.array([float(env.power_grid.current_signal)]),
        "base_power": np.array([float(env.power_grid.base_power)]),
        "artificial_ratio": np.array([float(env.power_grid.init_props.artificial_ratio)]),
        "max_power": np.array([float(env.cluster.max_power)]),
        "power": np.array([float(env.cluster.current_power_consumption)]),
        "solar": np.array([float(b[0].current_solar_gain)]),
    }
    if hasattr(env.power_grid, "time_since_last_interp"):
        st["t_since_interp"] = np.array([env.power_grid.time_since_last_interp], dtype=np.int64)
    return st


def load_controller_module(name: str):
    """Import ``app.core.agents.controllers.<name>`` WITHOUT running the package
    ``__init__`` (which drags in the trainers and ``pydantic.BaseSettings``, absent here)."""
    import importlib

    load()
    for pkg in ("app.core.agents", "app.core.agents.controllers"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(_SERVER, *pkg.split("."))]
            sys.modules[pkg] = m
    return importlib.import_module("app.core.agents.controllers." + name)
