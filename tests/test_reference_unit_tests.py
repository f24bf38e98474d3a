"""The reference's OWN unit tests for this path, restated against the CPU oracle (no GPU needed).

``server/v0/env/unit_tests_MA_DemandResponse.py`` (HVAC known answers :29-37, lock-out sequence
:39-70, solar-gain window :113-128, window/shading scaling :199-207, qualitative thermal behaviour
:209-553) and ``server/v0/monteCarlo/unit_tests_interp.py`` (grid-node exactness :67-90, index
round trip :92-107).  Same set-ups, same assertions, on ``oracle/np_oracle.py`` -- the restatement the
CUDA path is compared with in ``-m gpu``.  ``testDataPointsInter`` (:109-) needs the real table, which
is missing from the reference checkout; the synthetic table has no monotone structure to test.
"""
import datetime as dt

import numpy as np
import pytest

from oracle import np_oracle as orc
from oracle.config import INTERP_GRIDS, INTERP_KEYS, INTERP_SHAPE, synthetic_table

HOUSE = dict(Ua=2.18e02, Cm=3.45e06, Ca=9.08e05, Hm=2.84e03, window_area=7.175, shading_coeff=0.67)
HVAC = dict(cop=2.5, cooling_capacity=15000.0, latent_cooling_fraction=0.35, lockout_duration=12)
DT = 4


class House:
    """One ``SingleHouse`` (MA_DemandResponse.py:540-700) on the oracle's primitives."""

    def __init__(self, **over):
        p = dict(HOUSE, **over)
        self.__dict__.update(p)
        self.t_air = 20.0
        self.t_mass = 20.0
        self.on = False

    def step(self, od_temp, when):
        q_hvac = -HVAC["cooling_capacity"] / (1 + HVAC["latent_cooling_fraction"]) if self.on else 0.0
        qa = q_hvac + orc.solar_gain_scalar(when, self.window_area, self.shading_coeff)
        a, m = orc.thermal_update(np.float64(self.t_air), np.float64(self.t_mass), self.Ua, self.Ca, self.Cm, self.Hm,
                                  od_temp, qa, DT)
        self.t_air, self.t_mass = float(a), float(m)


def run_pair(h1, h2, od1, od2, t1, t2, checks):
    """step / 49 more / 950 more, with the reference's assertions after each stage"""
    step = dt.timedelta(seconds=DT)
    for n in (1, 49, 950):
        for _ in range(n):
            h1.step(od1, t1)
            h2.step(od2, t2)
            t1 += step
            t2 += step
        checks()


def test_hvac_consumption_and_heat_known_answers():          # :29-37
    cap, cop, lat = HVAC["cooling_capacity"], HVAC["cop"], HVAC["latent_cooling_fraction"]
    o = orc.NpOracle({"cluster_prop": {"nb_agents": 1, "house_prop": {"hvac_prop": HVAC}}}, 1)
    off = {"on": np.array([[False]]), "cap": np.array([[cap]])}
    on = {"on": np.array([[True]]), "cap": np.array([[cap]])}
    assert float(o.house_power(off).sum()) == 0.0
    assert float(o.house_power(on).sum()) == 15000 / 2.5 == cap / cop
    # heat extracted from the air (hvac.py:85-99), as the house update receives it
    a0, _ = orc.thermal_update(np.float64(20), np.float64(20), HOUSE["Ua"], HOUSE["Ca"], HOUSE["Cm"], HOUSE["Hm"], 20.0,
                               -15000 / (1 + 0.35), DT)
    a1, _ = orc.thermal_update(np.float64(20), np.float64(20), HOUSE["Ua"], HOUSE["Ca"], HOUSE["Cm"], HOUSE["Hm"], 20.0,
                               -cap / (1 + lat), DT)
    assert a0 == a1 < 20.0


def test_lockout_sequence():                                  # :39-70
    on, lock, sso = np.array([True]), np.array([False]), np.array([0])
    seq = []
    for action in (True, False, True, True, True, True):
        on, lock, sso = orc.hvac_fsm(on, lock, sso, np.array([action]), DT, HVAC["lockout_duration"])
        seq.append((bool(on[0]), bool(lock[0]), int(sso[0])))
    assert seq[0][:2] == (True, False)
    assert seq[1][:2] == (False, True)
    assert seq[2] == (False, True, 4)
    assert seq[3] == (False, True, 8)
    assert seq[4] == (True, False, 0)
    assert seq[5] == (True, False, 0)


def test_solar_gain_time_window():                            # :113-128
    g = lambda h, m: orc.solar_gain_scalar(dt.datetime(2021, 6, 15, h, m), HOUSE["window_area"], HOUSE["shading_coeff"])  # noqa: E731
    assert g(0, 0) == 0 and g(7, 29) == 0 and g(17, 31) == 0
    assert g(12, 0) > 0 and g(7, 31) > 0 and g(17, 29) > 0


def test_solar_gain_scales_with_window_and_shading():         # :199-207
    when = dt.datetime(2021, 6, 15, 12, 0)
    ori = orc.solar_gain_scalar(when, HOUSE["window_area"], HOUSE["shading_coeff"])
    mod = orc.solar_gain_scalar(when, HOUSE["window_area"] * 0.5, HOUSE["shading_coeff"] * 0.5)
    assert ori == mod * 4


MIDNIGHT = dt.datetime(2021, 6, 15, 0, 0)
MIDDAY = dt.datetime(2021, 6, 15, 12, 0)


def test_higher_initial_mass_temperature_heats_more():        # :209-249
    a, b = House(), House()
    a.t_mass, b.t_mass = 30.0, 22.0

    def checks():
        assert a.t_air > b.t_air and a.t_mass > b.t_mass
    run_pair(a, b, 25, 25, MIDNIGHT, MIDNIGHT, checks)


def test_higher_outdoor_temperature_heats_more():             # :251-292
    a, b = House(), House()

    def checks():
        assert a.t_air > b.t_air and a.t_mass > b.t_mass
    run_pair(a, b, 30, 22, MIDNIGHT, MIDNIGHT, checks)
    assert b.t_air > 20.0


@pytest.mark.parametrize("key,factor", [("Ua", 0.5), ("Ca", 2.0), ("Cm", 2.0)])   # :294-422
def test_wall_conductance_and_capacities(key, factor):
    a, b = House(), House(**{key: HOUSE[key] * factor})

    def checks():
        assert a.t_air > b.t_air and a.t_mass > b.t_mass
    run_pair(a, b, 30, 30, MIDNIGHT, MIDNIGHT, checks)
    assert b.t_air > 20.0


def test_mass_air_conductance():                              # :424-470
    a, b = House(), House(Hm=HOUSE["Hm"] / 2)
    a.t_mass = b.t_mass = 30.0

    def checks():
        assert a.t_air > b.t_air and a.t_mass < b.t_mass
    run_pair(a, b, 25, 25, MIDNIGHT, MIDNIGHT, checks)
    assert b.t_air > 20.0


def test_sun_effect():                                        # :472-515
    a, b = House(), House()

    def checks():
        assert a.t_air > b.t_air and a.t_mass > b.t_mass
    run_pair(a, b, 25, 25, MIDDAY, MIDNIGHT, checks)
    assert b.t_air > 20.0


def test_hvac_effect():                                       # :517-553
    a, b = House(), House()
    a.on = b.on = True

    def checks():
        assert a.t_air < 20.0 and a.t_mass < 20.0
    run_pair(a, b, 30, 30, MIDDAY, MIDDAY, checks)


# ---- unit_tests_interp.py -----------------------------------------------------------------------
def index_to_point(index):                                   # :15-29 (row-major, last key fastest)
    sub = np.unravel_index(index, INTERP_SHAPE)
    return {k: INTERP_GRIDS[k][i] for k, i in zip(INTERP_KEYS, sub)}


def point_to_index(point):                                   # :31-44
    idx = [list(INTERP_GRIDS[k]).index(point[k]) for k in INTERP_KEYS]
    return int(np.ravel_multi_index(idx, INTERP_SHAPE))


def interpolate(sub, point):
    default = dict(Ua=1.0, Cm=1.0, Ca=1.0, Hm=1.0)
    s = orc.interp_static_index(np.array([point["Ua_ratio"]]), np.array([point["Cm_ratio"]]), np.array([point["Ca_ratio"]]),
                                np.array([point["Hm_ratio"]]), np.array([point["HVAC_power"]]), default)
    return float(orc.interp_point(sub, s, np.array([point["air_temp"]]), np.array([point["mass_temp"]]),
                                  np.array([point["OD_temp"]]), np.array([point["hour"]]), np.array([point["date"]]))[0])


def test_interpolation_is_exact_on_grid_nodes():             # :67-90
    table = synthetic_table(2024)
    sub = orc.interp_sub_tables(table)
    flat = np.asarray(table).reshape(-1)
    n = flat.size
    assert n == 4_199_040
    for index in range(0, n, n // 50):
        assert interpolate(sub, index_to_point(index)) == flat[index], index


def test_index_point_round_trip():                           # :92-107
    air_0 = {"Ua_ratio": 1, "Cm_ratio": 1, "Ca_ratio": 1, "Hm_ratio": 1, "air_temp": -1, "mass_temp": 0, "OD_temp": 11,
             "HVAC_power": 15000, "hour": 11.0 * 3600, "date": 79}
    assert index_to_point(point_to_index(air_0)) == air_0
