"""drsim-b200: B200-native batched simulator of the demand-response environment step.

Public surface
--------------
``Environment``     drop-in for the reference's ``app.core.environment.environment.Environment``
                    (``reset()`` / ``step(action_dict)`` returning per-agent dicts).
``BatchedEnv``      thousands of independent replicas with zero-copy torch tensor views.
``norm_state_dict`` the reference's observation normaliser, served from device tensors.
``MADemandResponseEnv`` / ``norm_state_dict_v0``  the legacy gym-style env of ``server/v0`` (``v0.py``).
``ClientFeed``      the reference's UI feed payloads (``ClientManagerService``) from the simulator's tensors.
``ReferenceMetrics`` the reference's ``Metrics.update`` (metrics_service.py:108-157), literally, on the device.
``RolloutBuffer``   device-resident MA-PPO transition storage filled by ``BatchedEnv.collect`` (mappo.py:83-127).
``ShardedClusterEnv`` one very large cluster split by houses across the GPUs of a box.
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
    if name == "ReferenceMetrics":
        from . import metrics

        return metrics.ReferenceMetrics
    if name == "RolloutBuffer":
        from . import rollout

        return rollout.RolloutBuffer
    if name == "ShardedClusterEnv":
        from . import sharded

        return sharded.ShardedClusterEnv
    raise AttributeError(name)
