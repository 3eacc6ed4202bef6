"""Row N4, posterior state summaries: the oracle's restatement of ObservationModel::get_graph_state
(reference src/obsmodel.cc:113-170) against the reference's own, on the trajectories the reference itself simulated
(tests/golden/GRS_*.npz: harness mode BEEP_HARNESS_GRAPHS), and the library's host routine beep_graph_statistics against
a plain restatement of Output::get_statistic (output.cc:3089-3123)."""
import os

import numpy as np
import pytest

from beep_b200.model import ModelDesc
from oracle_lib import OracleState

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["GRS_C4", "GRS_C2", "GRS_B6", "GRS_B4"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_graph_state_matches_reference(name):
    d = dict(np.load(os.path.join(GOLD, name + ".npz")))
    m = ModelDesc.from_prefixed(d)
    assert m.graph_type.size > 0
    for k in range(d["theta"].shape[0]):
        for mode in ("det", "sto"):
            o = OracleState(m, stochastic=(mode == "sto"))
            o.set_param(d["theta"][k])
            o.init_from_particle(d[f"th{k}.{mode}.transnum"].reshape(m.traj_shape).astype(np.float64))
            ref = d[f"th{k}.{mode}.graph_state"]
            gs = o.graph_state()
            assert gs.shape == ref.shape
            assert np.abs(gs - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)


def test_graph_statistics_is_get_statistic():
    from beep_b200 import engine as E
    rng = np.random.default_rng(2)
    for P in (1, 2, 7, 200):
        gs = rng.gamma(2.0, 10.0, size=(P, 13))
        mean, lo, hi = E.graph_statistics(gs)
        for i in range(13):
            v = np.sort(gs[:, i])
            assert abs(mean[i] - gs[:, i].sum() / P) <= 1e-12 * abs(mean[i])
            if P >= 2:
                k = int((P - 1) * 0.025); f = (P - 1) * 0.025 - k
                assert lo[i] == v[k] * (1 - f) + v[k + 1] * f
                k = int((P - 1) * 0.975); f = (P - 1) * 0.975 - k
                assert hi[i] == v[k] * (1 - f) + v[k + 1] * f
            else:
                assert lo[i] == hi[i] == v[0]
