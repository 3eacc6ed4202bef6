"""The dense flavour of the area-mixing contraction (beep_b200/csrc/imap_dense.cuh): when every area mixes
with most others, Imap = M^T dinf (reference src/state.cc:204-216) of ALL particles of a step is one FP64
tensor-core GEMM (DMMA), followed by one step of the grid-spanning trajectory kernel.  Against the CPU oracle
(sequential sums): transition means to 1e-10 -- the deterministic piece of the north-star contract -- and
event counts equal (the sums differ by rounding only, far below what moves a draw); against the sparse-gather
flavour of the engine itself: the same trajectories."""
import os

import numpy as np
import pytest

from conftest import relerr
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu


def _engine(model, dense=None):
    from beep_b200 import engine as E
    if dense is not None:
        os.environ["BEEP_DENSE_M"] = "1" if dense else "0"
    try:
        return E.Engine(model)
    finally:
        os.environ.pop("BEEP_DENSE_M", None)


@pytest.mark.parametrize("A,G", [(96, 20), (130, 7)])
def test_dense_mixing_vs_oracle_and_gather(A, G):
    from beep_b200 import engine as E
    from beep_b200.synth import make_model, jitter_params
    m, theta = make_model(A=A, G=G, T=40, nneigh=A - 1, mean_pop=2e4)
    eng = _engine(m)
    assert eng.dense_mixing and eng.sync_scope == "GRID" and eng.chunk_tiles == 0      # chosen from the matrix itself
    P = 5
    th = jitter_params(theta, P, rel=0.05, seed=4)
    ids = np.arange(P, dtype=np.uint64) + 31
    flags = E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN
    b = eng.batch(P, flags)
    b.set_params(th)
    b.simulate(seed=12, stream_ids=ids)
    L = b.obs_loglik()
    k = 3
    st = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    st.set_param(th[k])
    st.simulate(seed=12, stream_id=int(ids[k]))
    tm = b.trajectory(k, E.TRAJ_TRANSMEAN)
    scale = float(np.max(np.abs(st.transmean)))
    assert relerr(tm, st.transmean, floor=1e-9 * scale) < 1e-10
    assert np.array_equal(b.trajectory(k).astype(np.int64), st.transnum.astype(np.int64))
    assert b.trajectory(k).sum() > 0
    assert relerr(L[k], st.obs_loglik()) < 1e-10
    # the sparse-gather flavour of the same engine
    eg = _engine(m, dense=False)
    assert not eg.dense_mixing
    g = eg.batch(P, flags)
    g.set_params(th)
    g.simulate(seed=12, stream_ids=ids)
    for s in range(P):
        assert np.array_equal(g.trajectory(s), b.trajectory(s))
    assert relerr(g.obs_loglik(), L) < 1e-12
    # model-based proposal and replay run through the same per-step GEMM
    prop = eng.batch(P, flags)
    th1 = th * (1 + 0.02 * (np.arange(th.shape[1]) % 5 == 2))
    prop.set_params(th1)
    prop.mbp(b, None, seed=13, stream_ids=ids)
    oi = st
    op = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    op.set_param(th1[k])
    op.mbp(oi, None, seed=13, stream_id=int(ids[k]))
    assert np.array_equal(prop.trajectory(k).astype(np.int64), op.transnum.astype(np.int64))
    # sectioned simulation (PMCMC) continues through the per-step launches
    sec = eng.batch(P, E.KEEP_TRANSNUM)
    sec.set_params(th)
    for s in range(eng.nsection):
        ti, tf = eng.section_range(s)
        sec.simulate(ti, tf, seed=12, stream_ids=ids)
    assert np.array_equal(sec.trajectory(k), b.trajectory(k))
