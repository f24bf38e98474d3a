"""The literal ``Metrics.update`` (metrics_service.py:108-157, precedence slips included) -- SURVEY 8f-3:
the NumPy restatement against values recorded from the REAL reference class (CPU), and the device kernel
``drsim_metrics_update`` against the same values (GPU)."""
import glob
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, GoldenCase

FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "metrics_*.json")))


def _before(case, t):
    """(t_air, signal, power, od_temp) of the observation BEFORE step t."""
    z, s0 = case.z, case.state0
    if t == 0:
        return np.asarray(s0["t_air"])[0], float(np.asarray(s0["signal"])[0]), float(np.asarray(s0["power"])[0]), \
            float(np.asarray(s0["od_temp"])[0])
    return z["t_air"][t - 1], float(z["signal"][t - 1]), float(z["power"][t - 1]), float(z["od_temp"][t - 1])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[8:-5] for f in FILES])
def test_restatement_matches_the_real_metrics_class(path):
    from oracle.metrics_ref import FIELDS, RefMetrics

    g = json.load(open(path))
    assert g["fields"] == FIELDS
    case = GoldenCase(g["case"])
    m = RefMetrics(case.N, g["start_stats_from"])
    tg = np.asarray(case.state0["target"])[0]
    for t in range(case.T):
        _, sig, p_old, od = _before(case, t)
        m.update(case.z["t_air"][t], tg, case.z["rewards"][t], sig, float(case.z["power"][t]), od, p_old, t)
        np.testing.assert_allclose(m.row(), g["per_step"][t], rtol=1e-13, atol=1e-12)
    rms = m.rms(case.T)
    for k, v in g["rms"].items():
        assert abs(rms[k] - v) <= 1e-12 * max(1.0, abs(v)), k


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[8:-5] for f in FILES])
def test_device_metrics_match_the_real_metrics_class(path):
    """The same trajectories replayed through the CUDA step (fp64 build); after every step the device accumulators
    of ``ReferenceMetrics`` (one launch of k_metrics_ref over the state planes) against the recorded values."""
    import torch

    from cuda_stepper import CudaStepper
    from marl_demandresponse_b200.metrics import ReferenceMetrics

    g = json.load(open(path))
    case = GoldenCase(g["case"])
    st = CudaStepper(case.env_prop, 1, "f64", "auto", table=case.table())
    st.set_state(case.state0)
    st.obs_vectors(case.comm(0))
    m = ReferenceMetrics(st.sim, start_stats_from=g["start_stats_from"])
    z = case.z
    uses_perlin = case.env_prop["power_grid_prop"]["signal_properties"]["mode"] == "perlin"
    for t in range(case.T):
        m.begin_step()
        ids = z["interp_ids"][t][None] if z["interp_ids"][t][0] >= 0 else None
        st.step(z["actions"][t][None], [z["od_noise"][t]], [z["perlin"][t]] if uses_perlin else None, ids)
        m.end_step(t)
        got = m.values()[0]
        want = np.asarray(g["per_step"][t])
        err, tol = np.abs(got - want), 1e-9 * np.maximum(1.0, np.abs(want))
        assert np.all(err <= tol), (t, [(f, float(a), float(b)) for f, a, b, e, l in zip(g["fields"], got, want, err, tol) if e > l])
    rms = m.rms(case.T)
    for k, v in g["rms"].items():
        assert abs(float(rms[k][0]) - v) <= 1e-9 * max(1.0, abs(v)), k
    torch.cuda.synchronize()
