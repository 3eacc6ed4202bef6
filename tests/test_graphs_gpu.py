"""Row N4 on the device: beep_graph_state (beep_b200/csrc/graphs.cuh) through the C ABI -- the series behind the
reference's Posterior/state files, computed from the batch's resident trajectory -- against the reference's own
ObservationModel::get_graph_state on the same trajectories (tests/golden/GRS_*.npz; 1e-10: the device sums in a tree)
and against the oracle on trajectories the engine simulated itself."""
import os

import numpy as np
import pytest

from beep_b200 import engine as E
from beep_b200.model import ModelDesc
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["GRS_C4", "GRS_C2", "GRS_B6", "GRS_B4"]


@pytest.mark.parametrize("name", CASES)
def test_graph_state_of_reference_trajectories(name):
    d = dict(np.load(os.path.join(GOLD, name + ".npz")))
    m = ModelDesc.from_prefixed(d)
    eng = E.Engine(m, 0)
    K = d["theta"].shape[0]
    bt = eng.batch(K, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN | E.KEEP_POP)
    for k in range(K):
        bt.import_particle(k, d["theta"][k], d[f"th{k}.sto.transnum"].reshape(m.traj_shape).astype(np.float64))
    bt.replay()
    gs = bt.graph_state()
    for k in range(K):
        ref = d[f"th{k}.sto.graph_state"]
        assert np.abs(gs[k] - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1.0), (name, k)


def test_graph_state_and_summary_of_simulated_batch():
    d = dict(np.load(os.path.join(GOLD, "GRS_C4.npz")))
    m = ModelDesc.from_prefixed(d)
    eng = E.Engine(m, 0)
    P = 64
    th = np.repeat(d["theta"][:1], P, axis=0)
    bt = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_POP)
    bt.set_params(th)
    bt.simulate(seed=9)
    gs = bt.graph_state()
    for k in (0, 63):
        o = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
        o.set_param(th[k])
        o.simulate(seed=9, stream_id=k)
        ref = o.graph_state()
        assert np.abs(gs[k] - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1.0)
    mean, lo, hi = E.graph_statistics(gs)
    assert np.all(lo <= hi) and np.any(hi > lo) and np.allclose(mean, gs.mean(axis=0), rtol=1e-12)


def test_graph_state_needs_pop_when_a_graph_plots_compartments():
    d = dict(np.load(os.path.join(GOLD, "GRS_C4.npz")))
    m = ModelDesc.from_prefixed(d)
    eng = E.Engine(m, 0)
    bt = eng.batch(2, E.KEEP_TRANSNUM)
    bt.set_params(d["theta"][:2])
    bt.simulate(seed=1)
    with pytest.raises(E.BeepError):
        bt.graph_state()


def test_marginal_graph_points_are_observation_values():
    """GRAPH_MARGINAL (obsmodel.cc:159-165): a point is the value of its observation.  No example that loads at this commit
    has one (EX_B1 aborts in the reference itself), so a marginal graph over five observations is appended to a fixture's
    description; device, oracle and beep_obs_values must agree."""
    d = dict(np.load(os.path.join(GOLD, "GRS_B4.npz")))
    m0 = ModelDesc.from_prefixed(d)
    arrays = dict(m0.arrays)
    obs = np.array([0, 3, 7, 11, 20], dtype=np.int32)
    ng = m0.graph_type.size
    arrays["graph_type"] = np.append(m0.graph_type, 1).astype(np.int32)
    arrays["graph_table"] = np.append(m0.graph_table, m0.obs_table[obs[0]]).astype(np.int32)
    arrays["graph_factor"] = np.append(m0.graph_factor, 1.0)
    arrays["graph_factor_spline"] = np.append(m0.graph_factor_spline, -1).astype(np.int32)
    arrays["graph_area_off"] = np.append(m0.graph_area_off, m0.graph_area_off[-1]).astype(np.int32)
    arrays["graph_dp_off"] = np.append(m0.graph_dp_off, m0.graph_dp_off[-1]).astype(np.int32)
    arrays["graph_point_off"] = np.append(m0.graph_point_off, m0.graph_point_off[-1] + obs.size).astype(np.int32)
    arrays["graph_point_obs"] = np.append(m0.graph_point_obs, obs).astype(np.int32)
    m = ModelDesc(arrays)
    eng = E.Engine(m, 0)
    bt = eng.batch(3, E.KEEP_TRANSNUM | E.KEEP_POP)
    bt.set_params(np.repeat(d["theta"][:1], 3, axis=0))
    bt.simulate(seed=4)
    gs = bt.graph_state()
    ov = bt.obs_values()
    assert gs.shape[1] == d["th0.sto.graph_state"].size + obs.size
    assert np.array_equal(gs[:, -obs.size:], ov[:, obs])
    o = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    o.set_param(d["theta"][0])
    o.simulate(seed=4, stream_id=1)
    ref = o.graph_state()
    assert np.abs(gs[1] - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1.0)
    assert ng + 1 == m.graph_type.size
