"""drsim-b200: B200-native batched simulator of the demand-response environment step.

Public surface
--------------
``Environment``     drop-in for the reference's ``app.core.environment.environment.Environment``
                    (``reset()`` / ``step(action_dict)`` returning per-agent dicts).
``BatchedEnv``      thousands of independent replicas with zero-copy torch tensor views.
``norm_state_dict`` the reference's observation normaliser, served from device tensors.
``MADemandResponseEnv`` / ``norm_state_dict_v0``  the legacy gym-style env of ``server/v0`` (``v0.py``).
``ClientFeed``      the reference's UI feed payloads (``ClientManagerService``) from the simulator's tensors.
``DrSim``           thin owner of the C handle (``include/drsim.h``).

The CUDA extension (``libdrsim.so``, sm_100a) is mandatory; there is no CPU fallback.
"""
from .core import DrSim, flatten_config, from_epoch, to_epoch  # noqa: F401
from .properties import EnvironmentProperties, as_props  # noqa: F401


def __getattr__(name):
    if name in ("Environment", "norm_state_dict"):
        from . import environment

        return getattr(environment, name)
    if name in ("MADemandResponseEnv", "norm_state_dict_v0", "props_from_v0"):
        from . import v0

        return getattr(v0, name)
    if name in ("BatchedEnv", "synthetic_state"):
        from . import batched

        return getattr(batched, name)
    if name in ("ClientFeed", "DESCRIPTION_KEYS"):
        from . import ui_feed

        return getattr(ui_feed, name)
    raise AttributeError(name)
