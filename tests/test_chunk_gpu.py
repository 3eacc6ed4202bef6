"""The chunk kernel (beep_b200/csrc/traj_chunk.cuh: one warp per chunk of 4 or 8 tiles, int32
populations, pooled draws) against the CPU oracle -- bit-exact event counts, 1e-10 on doubles -- in
all three modes (simulate / model-based proposal / replay), both chunk sizes, with ragged last
chunks, dense models (a list flush per tile, rejection samplers) and time-varying age matrices;
and against the one-tile-per-warp kernel it replaces (same bits)."""
import os

import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu


def _engine(model, nt=None, chunk=True):
    from beep_b200 import engine as E
    os.environ["BEEP_FORCE_SYNC"] = "2"
    if nt:
        os.environ["BEEP_CHUNK_NT"] = str(nt)
    if not chunk:
        os.environ["BEEP_NO_CHUNK"] = "1"
    try:
        return E.Engine(model)
    finally:
        for k in ("BEEP_FORCE_SYNC", "BEEP_CHUNK_NT", "BEEP_NO_CHUNK"):
            os.environ.pop(k, None)


def _models():
    from beep_b200.synth import make_model
    return {
        "sparse37": lambda: make_model(A=37, G=20, T=90),                       # ragged last chunk, sparse epidemic (pooled flush)
        "dense13": lambda: make_model(A=13, G=20, T=70, mean_pop=3e6),          # large means: rejection samplers, a flush per tile
        "ages6": lambda: make_model(A=10, G=6, T=60, mean_pop=2e4),             # G does not divide 32: pairs straddle passes
    }


@pytest.mark.parametrize("nt", [4, 8])
@pytest.mark.parametrize("name", ["sparse37", "dense13", "ages6"])
def test_chunk_kernel_vs_oracle(name, nt):
    from beep_b200 import engine as E
    from beep_b200.synth import jitter_params
    m, theta = _models()[name]()
    eng = _engine(m, nt)
    assert eng.sync_scope == "GRID" and eng.chunk_tiles == nt
    P = 3
    th0 = jitter_params(theta, P, rel=0.05, seed=8)
    th1 = th0 * (1 + 0.02 * (np.arange(th0.shape[1]) % 5 == 2))
    ids = np.array([14, 3, 9], dtype=np.uint64)
    flags = E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN
    init = eng.batch(P, flags)
    init.set_params(th0)
    init.simulate(seed=33, stream_ids=ids)
    L0, tab0 = init.obs_loglik(with_tables=True)
    ov0 = init.obs_values()
    Li0 = init.latent_loglik()
    k = 1
    oi = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    oi.set_param(th0[k])
    oi.simulate(seed=33, stream_id=int(ids[k]))
    tn = init.trajectory(k)
    assert np.array_equal(tn.astype(np.int64), oi.transnum.astype(np.int64))
    assert tn.sum() > 0
    scale = float(np.max(np.abs(oi.transmean)))
    assert relerr(init.trajectory(k, E.TRAJ_TRANSMEAN), oi.transmean, floor=1e-9 * scale) < 1e-10
    assert relerr(ov0[k], oi.obs_values(), floor=1e-9) < 1e-10
    assert relerr(L0[k], oi.obs_loglik()) < 1e-10
    assert relerr(tab0[k], oi.EF_datatable(), floor=1e-9) < 1e-10
    assert relerr(Li0[k], oi.latent_loglik()) < 1e-10
    assert np.array_equal(init.pop(k), oi.pop[m.T - 1])
    # the trajectory-discarding instantiation gives the same observation sums
    nk = eng.batch(P, 0)
    nk.set_params(th0)
    nk.simulate(seed=33, stream_ids=ids)
    assert np.array_equal(nk.obs_values(), ov0)
    assert np.array_equal(nk.pop(k), init.pop(k))
    # sectioned simulation continues from the resident integer population
    sec = eng.batch(P, E.KEEP_TRANSNUM)
    sec.set_params(th0)
    for s in range(eng.nsection):
        ti, tf = eng.section_range(s)
        sec.simulate(ti, tf, seed=33, stream_ids=ids)
    assert np.array_equal(sec.trajectory(k), tn)
    # model-based proposal (pure and with a SIMU mask), then replay under other parameters
    prop = eng.batch(P, flags)
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
        assert np.array_equal(prop.pop(k), op.pop[m.T - 1])
    for s in range(P):
        prop.import_particle(s, th1[s], init.trajectory(s))
    prop.replay()
    o3 = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
    o3.set_param(th1[k])
    o3.init_from_particle(init.trajectory(k))
    scale = float(np.max(np.abs(o3.transmean)))
    assert relerr(prop.trajectory(k, E.TRAJ_TRANSMEAN), o3.transmean, floor=1e-9 * scale) < 1e-10
    assert relerr(prop.latent_loglik()[k], o3.latent_loglik()) < 1e-10
    assert relerr(prop.obs_values()[k], o3.obs_values(), floor=1e-9) < 1e-10
    assert np.array_equal(prop.pop(k), o3.pop[m.T - 1])
    # accept / permute move integer populations like the FP64 ones
    before = init.pop(0)
    init.accept_from(prop, np.array([0, 1, 0], dtype=np.uint8))
    assert np.array_equal(init.pop(1), prop.pop(1)) and np.array_equal(init.pop(0), before)
    init.permute(np.array([2, E.KEEP, 1], dtype=np.uint32), E.PERMUTE_SECTION_STATE)
    assert np.array_equal(init.pop(2), prop.pop(1))
    pa = init.pop_all()
    init.set_pop_all(pa[::-1].copy(), m.T)
    assert np.array_equal(init.pop(0), pa[2])


@pytest.mark.parametrize("name", ["sparse37", "dense13"])
def test_chunk_kernel_equals_tile_kernel(name):
    """Same bits from the kernel it replaces (k_traj<SIM, SYNC_GRID>, FP64 populations)."""
    from beep_b200 import engine as E
    from beep_b200.synth import jitter_params
    m, theta = _models()[name]()
    P = 4
    th = jitter_params(theta, P, rel=0.05, seed=2)
    outs = []
    for chunk in (True, False):
        eng = _engine(m, 8, chunk)
        assert (eng.chunk_tiles != 0) == chunk
        b = eng.batch(P, E.KEEP_TRANSNUM)
        b.set_params(th)
        b.simulate(seed=5)
        outs.append(([b.trajectory(k) for k in range(P)], b.obs_values(), b.obs_loglik(), [b.pop(k) for k in range(P)]))
    for k in range(P):
        assert np.array_equal(outs[0][0][k], outs[1][0][k])
        assert np.array_equal(outs[0][3][k], outs[1][3][k])
    assert relerr(outs[0][1], outs[1][1], floor=1e-9) < 1e-12
    assert relerr(outs[0][2], outs[1][2]) < 1e-12


@pytest.mark.parametrize("name", ["EX_E2", "EX_C4", "EX_B4"])
def test_chunk_kernel_on_reference_models(name):
    """Worked examples of the reference that have the chunk kernel's structure (ages, one strain),
    incl. a time-varying age matrix (EX_E2), forced onto the grid-spanning flavour."""
    from beep_b200 import engine as E
    from beep_b200.model import ModelDesc
    m, d = load_golden(name)
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    ms = ModelDesc(arrays)
    eng = _engine(ms, 4)
    if eng.chunk_tiles == 0:
        pytest.skip("model does not have the chunk kernel's structure")
    P = 3
    theta = np.repeat(d["theta"][:1], P, axis=0)
    b = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(theta)
    b.simulate(seed=3)
    st = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    st.set_param(theta[0])
    st.simulate(seed=3, stream_id=2)
    assert np.array_equal(b.trajectory(2).astype(np.int64), st.transnum.astype(np.int64))
    assert relerr(b.obs_values()[2], st.obs_values(), floor=1e-9) < 1e-10
    assert relerr(b.obs_loglik()[2], st.obs_loglik()) < 1e-10
