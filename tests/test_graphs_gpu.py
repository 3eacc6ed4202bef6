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
