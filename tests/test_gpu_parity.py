"""GPU parity tests proper: the CUDA path, called through the C ABI, against the golden
trajectories recorded from the reference and against the NumPy oracle.

Tolerances (BASELINE.json north_star): discrete state (on / lockout / seconds_since_off, epoch)
bit-exact; continuous quantities within rtol 1e-5 (fp32 build) or 1e-12 (fp64 build) of their
natural scale (temperatures 20 degC, powers 6 kW x N; normalised observations and rewards 1 in fp32, 4 = 20 degC / 5
and 40 = 2 x 20 degC in fp64 -- see golden_util.scales_for).  DRSIM_PARITY_OUT=<file> collects the worst errors per case.
"""
import numpy as np
import pytest

from golden_util import GoldenCase, case_names, record_worst, replay

pytestmark = pytest.mark.gpu

RTOL = {"f32": 1e-5, "f64": 1e-12}


def _stepper(case, precision, path, **kw):
    from cuda_stepper import CudaStepper

    if path == "fused" and case.env_prop["power_grid_prop"]["base_power_props"]["mode"] == "interpolation":
        path = "auto"  # the interpolator re-evaluation always runs on the general path
    return CudaStepper(case.env_prop, 1, precision, path, table=case.table(), **kw)


@pytest.mark.parametrize("path", ["fused", "split"])
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", case_names())
def test_golden_trajectory(name, precision, path):
    case = GoldenCase(name)
    st = _stepper(case, precision, path)
    worst = replay(case, st, rtol=RTOL[precision], precision=precision)
    print(name, precision, path, {k: f"{v:.2e}" for k, v in worst.items()})
    record_worst(name, precision, path, worst)


@pytest.mark.parametrize("name", ["c1_default_n10_bangbang", "random_n24_sinus_commonL2"])
def test_device_pointer_api_equals_host_api(name):
    case = GoldenCase(name)
    a = _stepper(case, "f32", "auto", host_api=True)
    b = _stepper(case, "f32", "auto", host_api=False)
    wa = replay(case, a, rtol=1e-5)
    wb = replay(case, b, rtol=1e-5)
    assert wa == wb


def test_per_step_error_fp32_is_at_rounding_level():
    """Re-anchoring the fp32 state on the golden trajectory every step isolates the single-step
    error of the difference-form update: <= 1 ulp of a 20 degC temperature (1.9e-6)."""
    case = GoldenCase("random_n24_sinus_commonL2")
    st = _stepper(case, "f32", "auto")
    worst = replay(case, st, rtol=1e-5, reinject_every=1, check_obs=False)
    assert worst["t_air"] < 4e-6 and worst["t_mass"] < 4e-6, worst
