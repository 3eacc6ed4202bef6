"""Parity at BASELINE.json's full synthetic size (configs[4]: 10 000 areas x 20 ages x 12 compartments
x 365 steps) through properties that do not need the oracle at that size, plus a direct oracle
comparison at the largest size the oracle finishes in seconds.

The grid-spanning kernel hands tiles out dynamically (per-particle step scheduler), so which warp
works which tile in which order differs from launch to launch and with the batch's composition.
Every cell owns its Philox stream and every sum has a fixed order, so NONE of that may show in
the results:
  * determinism: the same launch twice gives the same bits;
  * batch-composition invariance: a particle (parameters, stream id) gives the same bits whatever
    the batch around it and whatever its slot -- the keyed-stream contract;
  * conservation: transitions only move individuals between compartments (state.cc:460-485), so the
    population of every (area, dp) summed over compartments is what it was at t = 0 (single
    compartments may dip below zero, as in the reference: a Poisson count is not capped by its source);
  * the observation log-likelihood is finite and the tau-leap did something (the epidemic moved).
"""
import numpy as np
import pytest

from conftest import relerr
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu


def _run(eng, th, ids, seed):
    b = eng.batch(len(ids), 0)
    b.set_params(th)
    b.simulate(seed=seed, stream_ids=np.asarray(ids, dtype=np.uint64))
    L = b.obs_loglik()
    obs = b.obs_values()
    pops = [b.pop(k) for k in range(len(ids))]
    b.close()
    return L, obs, pops


def test_full_size_properties():
    from beep_b200 import engine as E
    from beep_b200.synth import make_model, jitter_params
    m, theta = make_model(A=10000, G=20, T=365)
    assert (m.A, m.D, m.C, m.T) == (10000, 20, 12, 365)
    eng = E.Engine(m)
    assert eng.sync_scope == "GRID"
    th = jitter_params(theta, 6, rel=0.05, seed=5)
    ids = [11, 3, 7, 5, 2, 9]
    L6, obs6, pop6 = _run(eng, th, ids, seed=21)
    # determinism
    L6b, obs6b, pop6b = _run(eng, th, ids, seed=21)
    assert np.array_equal(L6, L6b) and np.array_equal(obs6, obs6b)
    for a, b in zip(pop6, pop6b):
        assert np.array_equal(a, b)
    # batch-composition invariance: two of the particles alone, in other slots
    sel = [4, 1]
    L2, obs2, pop2 = _run(eng, th[sel], [ids[k] for k in sel], seed=21)
    for j, k in enumerate(sel):
        assert L2[j] == L6[k]
        assert np.array_equal(obs2[j], obs6[k])
        assert np.array_equal(pop2[j], pop6[k])
    # another seed is another trajectory
    L6c, _, _ = _run(eng, th, ids, seed=22)
    assert not np.array_equal(L6, L6c)
    # conservation
    pop0 = np.asarray(m.arrays["pop_init"], dtype=np.float64).reshape(m.A, m.C, m.D)
    for p in pop6:
        assert np.array_equal(p.sum(axis=1), pop0.sum(axis=1))
        assert not np.array_equal(p, pop0)
    assert np.all(np.isfinite(L6))


def test_largest_oracle_size_bit_exact():
    """1 000 areas x 20 ages x 365 steps through the grid-spanning kernel against the oracle:
    final populations and every observation value (weekly sums of event counts per area) equal
    bit for bit, log-likelihood to 1e-10."""
    from beep_b200 import engine as E
    from beep_b200.synth import make_model, jitter_params
    m, theta = make_model(A=1000, G=20, T=365)
    import os
    os.environ["BEEP_FORCE_SYNC"] = "2"
    try:
        eng = E.Engine(m)
    finally:
        os.environ.pop("BEEP_FORCE_SYNC", None)
    assert eng.sync_scope == "GRID"
    th = jitter_params(theta, 3, rel=0.05, seed=8)
    ids = [4, 1, 6]
    L, obs, pops = _run(eng, th, ids, seed=33)
    k = 2
    st = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    st.set_param(th[k])
    st.simulate(seed=33, stream_id=ids[k])
    assert np.array_equal(obs[k], st.obs_values())
    assert np.array_equal(pops[k], st.pop[m.T - 1])
    assert relerr(L[k], st.obs_loglik()) < 1e-10


def test_largest_oracle_size_mbp_replay_bit_exact():
    """The grid-spanning MBP and replay kernels (k_traj<MBP,GRID>, k_traj<REPLAY,GRID> -- what
    `bench.py --mode mbp` times) at 1 000 areas against the oracle: the FULL proposed transnum
    (mbp.cc:252-330) bit for bit, transition means of the replay (state.cc:294-322) to 1e-10."""
    from beep_b200 import engine as E
    from beep_b200.synth import make_model, jitter_params
    import os
    m, theta = make_model(A=1000, G=20, T=120)
    os.environ["BEEP_FORCE_SYNC"] = "2"
    try:
        eng = E.Engine(m)
    finally:
        os.environ.pop("BEEP_FORCE_SYNC", None)
    assert eng.sync_scope == "GRID"
    P = 3
    th0 = jitter_params(theta, P, rel=0.05, seed=8)
    th1 = th0 * (1 + 0.01 * (np.arange(th0.shape[1]) % 7 == 3))
    ids = np.array([4, 1, 6], dtype=np.uint64)
    flags = E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN
    init = eng.batch(P, flags)
    prop = eng.batch(P, flags)
    init.set_params(th0)
    init.simulate(seed=33, stream_ids=ids)
    k = 1
    oi = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    oi.set_param(th0[k])
    oi.simulate(seed=33, stream_id=int(ids[k]))
    assert np.array_equal(init.trajectory(k).astype(np.int64), oi.transnum.astype(np.int64))
    mask = np.zeros((m.T, m.A), dtype=np.uint8)
    mask[m.T // 2:, 1::3] = 1
    for msk, seed in ((None, 34), (mask, 35)):
        prop.set_params(th1)
        prop.mbp(init, msk, seed=seed, stream_ids=ids)
        op = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
        op.set_param(th1[k])
        op.mbp(oi, msk, seed=seed, stream_id=int(ids[k]))
        assert np.array_equal(prop.trajectory(k).astype(np.int64), op.transnum.astype(np.int64))
        scale = float(np.max(np.abs(op.transmean)))
        assert relerr(prop.trajectory(k, E.TRAJ_TRANSMEAN), op.transmean, floor=1e-9 * scale) < 1e-10
        assert relerr(prop.obs_loglik()[k], op.obs_loglik()) < 1e-10
    # replay of the initial events under the proposal's parameters
    for s in range(P):
        prop.import_particle(s, th1[s], init.trajectory(s))
    prop.replay()
    o3 = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    o3.set_param(th1[k])
    o3.init_from_particle(init.trajectory(k))
    scale = float(np.max(np.abs(o3.transmean)))
    assert relerr(prop.trajectory(k, E.TRAJ_TRANSMEAN), o3.transmean, floor=1e-9 * scale) < 1e-10
    assert relerr(prop.latent_loglik()[k], o3.latent_loglik()) < 1e-10
    assert np.array_equal(prop.pop(k), o3.pop[m.T - 1])
