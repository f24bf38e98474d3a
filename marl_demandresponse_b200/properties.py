"""Configuration models of the drop-in environment.

Field names, nesting and defaults mirror the reference's pydantic models so that the
``env_prop`` block of ``server/app/core/config/MARLconfig.json`` parses unchanged and so that
``utils/norm.py:norm_state_dict(obs, env.init_props)`` keeps working on our ``init_props``:

* ``core/environment/environment_properties.py`` (``HvacProperties`` :57, ``BuildingProperties``
  :162, ``RewardProperties`` :229, ``StateProperties`` :261, ``MessageProperties`` :294,
  ``ClusterPropreties`` :313 [sic], ``EnvironmentProperties`` :339)
* ``core/environment/cluster/cluster_properties.py`` (``TemperatureProperties`` :4,
  ``AgentsCommunicationProperties`` :31)
* ``core/environment/power_grid/power_grid_properties.py`` (:6, :30, :56)

The environment itself also accepts the reference's own model instances or a plain nested
dict (see :func:`as_dict`).
"""
from __future__ import annotations

import datetime as _dt
from typing import Any, List, Literal

from pydantic import BaseModel, Field


class HvacNoiseProperties(BaseModel):
    std_latent_cooling_fraction: float = 0.05
    factor_COP_low: float = 0.95
    factor_COP_high: float = 1.05
    factor_cooling_capacity_low: float = 0.9
    factor_cooling_capacity_high: float = 1.1
    lockout_noise: int = 0
    cooling_capacity_list: List[int] = [12500, 15000, 17500]


class HvacProperties(BaseModel):
    cop: float = Field(default=2.5, gt=0)
    cooling_capacity: float = Field(default=15000.0, gt=0)
    latent_cooling_fraction: float = Field(default=0.35, gt=0, lt=1)
    lockout_duration: int = 40
    noise_prop: HvacNoiseProperties = HvacNoiseProperties()

    @property
    def max_consumption(self) -> float:
        return self.cooling_capacity / self.cop


class BuildingNoiseProperties(BaseModel):
    std_start_temp: float = 3.0
    std_target_temp: float = 1.0
    factor_thermo_low: float = 0.9
    factor_thermo_high: float = 1.1


class ThermalProperties(BaseModel):
    Ua: float = 2.18e02
    Ca: float = 9.08e05
    Hm: float = 2.84e03
    Cm: float = 3.45e06


class BuildingProperties(ThermalProperties):
    target_temp: float = 20.0
    deadband: float = 0.0
    init_air_temp: float = 20.0
    init_mass_temp: float = 20.0
    solar_gain: bool = True
    window_area: float = 7.175
    shading_coeff: float = 0.67
    noise_prop: BuildingNoiseProperties = BuildingNoiseProperties()
    hvac_prop: HvacProperties = HvacProperties()


class PenaltyProperties(BaseModel):
    mode: Literal["common_L2", "individual_L2", "common_max_error", "mixture"] = "individual_L2"
    alpha_ind_l2: float = 1.0
    alpha_common_l2: float = 1.0
    alpha_common_max: float = 0.0


class RewardProperties(BaseModel):
    alpha_temp: float = 1.0
    alpha_sig: float = 1.0
    norm_reg_sig: int = 7500
    penalty_props: PenaltyProperties = PenaltyProperties()
    sig_penalty_mode: Literal["common_L2"] = "common_L2"


class StateProperties(BaseModel):
    hour: bool = False
    day: bool = False
    solar_gain: bool = False
    thermal: bool = False
    hvac: bool = False


class MessageProperties(BaseModel):
    thermal: bool = False
    hvac: bool = False


class TemperatureProperties(BaseModel):
    day_temp: float = 26.0
    night_temp: float = 20.0
    temp_std: float = 1.0
    random_phase_offset: bool = False
    phase: float = 0.0


class AgentsCommunicationProperties(BaseModel):
    mode: str = "neighbours"
    row_size: int = 5
    max_communication_distance: int = 2
    max_nb_agents_communication: int = 10


class ClusterPropreties(BaseModel):  # spelling follows the reference
    nb_agents: int = 1000
    nb_agents_comm: int = 10
    agents_comm_prop: AgentsCommunicationProperties = AgentsCommunicationProperties()
    message_prop: MessageProperties = MessageProperties()
    house_prop: BuildingProperties = BuildingProperties()


class SignalProperties(BaseModel):
    mode: str = "perlin"
    amplitude_ratios: List[float] = [0.1, 0.3]
    amplitude_per_hvac: int = 6000
    nb_octaves: int = 5
    octaves_step: int = 5
    period: int = 300
    periods: List[int] = [400, 1200]


class BasePowerProperties(BaseModel):
    mode: str = "constant"
    avg_power_per_hvac: int = 4200
    init_signal_per_hvac: int = 910
    path_datafile: str = "./monteCarlo/mergedGridSearchResultFinal.npy"
    path_parameter_dict: str = "./monteCarlo/interp_parameters_dict.json"
    path_dict_keys: str = "./monteCarlo/interp_dict_keys.csv"
    interp_update_period: int = 300
    interp_nb_agents: int = 100


class PowerGridProperties(BaseModel):
    artificial_signal_ratio_range: int = 1
    base_power_props: BasePowerProperties = BasePowerProperties()
    signal_properties: SignalProperties = SignalProperties()
    artificial_ratio: float = 1.0


class EnvironmentProperties(BaseModel):
    start_datetime: _dt.datetime = _dt.datetime(2021, 1, 1, 0, 0, 0)
    start_datetime_mode: Literal["fixed", "random"] = "random"
    time_step: _dt.timedelta = _dt.timedelta(0, 4)
    temp_prop: TemperatureProperties = TemperatureProperties()
    state_prop: StateProperties = StateProperties()
    reward_prop: RewardProperties = RewardProperties()
    cluster_prop: ClusterPropreties = ClusterPropreties()
    power_grid_prop: PowerGridProperties = PowerGridProperties()


def as_props(obj: Any) -> EnvironmentProperties:
    """Coerce a dict / reference pydantic model / our model into ``EnvironmentProperties``."""
    if isinstance(obj, EnvironmentProperties):
        return obj.model_copy(deep=True)
    if obj is None:
        return EnvironmentProperties()
    if isinstance(obj, dict):
        return EnvironmentProperties(**obj)
    for dump in ("model_dump", "dict"):
        if hasattr(obj, dump):
            return EnvironmentProperties(**getattr(obj, dump)())
    raise TypeError(f"cannot interpret {type(obj)!r} as EnvironmentProperties")
