"""Parity of the CUDA engine (through the C-ABI) against the CPU oracle and the
committed golden vectors of the reference.  Every test here needs a GPU.

Tolerances (BASELINE.json north_star):
  * deterministic pieces (derived tables, force of infection / transition means,
    observation log-likelihood on a fixed trajectory, latent likelihood): 1e-10
    relative in FP64;
  * integer event counts under the shared counter-based stream: bit-exact.
The oracle runs in KEYED mode (one Philox stream per cell) with the infectivity
map derived from the population each step -- the engine's scheme; the oracle's
SEQ mode is what test_oracle_golden.py pins against the unmodified reference.
"""
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden, relerr
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu

RTOL = 1e-10
NAMES = golden_names()


def _engine(model, force_sync=None):
    from beep_b200 import engine
    if force_sync is None:
        os.environ.pop("BEEP_FORCE_SYNC", None)
    else:
        os.environ["BEEP_FORCE_SYNC"] = str(force_sync)
    try:
        return engine.Engine(model)
    finally:
        os.environ.pop("BEEP_FORCE_SYNC", None)


def _variant(model, **over):
    from beep_b200.model import ModelDesc
    arrays = dict(model.arrays)
    arrays.update(over)
    return ModelDesc(arrays)


def test_sampler_bit_exact():
    """Device Poisson / binomial == host ones, draw for draw (include/beep_rng.h)."""
    from beep_b200 import engine
    import oracle_lib
    import ctypes
    L = oracle_lib.lib()
    rng = np.random.default_rng(1)
    lam = np.concatenate([10 ** rng.uniform(-6, 6, 20000), rng.uniform(9.5, 10.5, 2000), [0.0, 1e-300, 9.999999, 10.0, 1e9]])
    host = np.zeros(lam.size, dtype=np.int32)
    L.oracle_poisson_many(123, 456, lam.ctypes.data, lam.size, host.ctypes.data)
    dev = engine.test_poisson(123, 456, lam)
    assert np.array_equal(host, dev)
    # sanity of the distribution itself
    big = lam > 100
    z = (dev[big] - lam[big]) / np.sqrt(lam[big])
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05
    n = np.concatenate([rng.integers(0, 50, 10000), rng.integers(0, 2_000_000, 10000)]).astype(np.int32)
    p = np.concatenate([rng.uniform(0, 1, 15000), 10 ** rng.uniform(-8, 0, 5000)])
    hostb = np.zeros(n.size, dtype=np.int32)
    L.oracle_binomial_many(7, 8, p.ctypes.data, n.ctypes.data, n.size, hostb.ctypes.data)
    devb = engine.test_binomial(7, 8, p, n)
    assert np.array_equal(hostb, devb)
    assert np.all(devb >= 0) and np.all(devb <= n)


@pytest.mark.parametrize("name", NAMES)
def test_set_param(name):
    """K0 vs the reference's State::set_param (golden) and vs the oracle."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    eng = _engine(m)
    nth = d["theta"].shape[0]
    b = eng.batch(nth)
    b.set_params(d["theta"])
    st = OracleState(m)
    for k in range(nth):
        p = f"th{k}."
        tr = b.derived(k, E.DERIVED_TRANSRATE)
        assert relerr(tr[m.trans_inf == 0].ravel(), d[p + "transrate"]) < RTOL
        assert relerr(b.derived(k, E.DERIVED_DISC_SPLINE).ravel(), d[p + "disc_spline"]) < RTOL
        assert relerr(b.derived(k, E.DERIVED_SUSCEPTIBILITY), d[p + "susceptibility"]) < RTOL
        assert relerr(b.derived(k, E.DERIVED_AREAFACTOR).ravel(), d[p + "areafactor"]) < RTOL
        assert relerr(b.derived(k, E.DERIVED_BETA).ravel(), d[p + "beta"]) < RTOL
        # time-varying age matrix through its factorisation N0[a][b] g_a g_b
        g = b.derived(k, E.DERIVED_AGEFACTOR)
        T = m.T
        N0 = m.N0.reshape(m.G, m.G)
        nt3 = np.concatenate([(N0 * np.outer(g[s], g[s])).ravel() for s in (0, T // 2, T - 1)])
        assert relerr(nt3, d[p + "Ntime3"]) < RTOL
        # against the oracle the derived tables are BIT-identical (same operations in the same order,
        # no FMA contraction on either side): this is what makes the integer draws bit-exact
        st.set_param(d["theta"][k])
        assert np.array_equal(b.derived(k, E.DERIVED_TRANSRATE), st.transrate)
        assert np.array_equal(b.derived(k, E.DERIVED_DISC_SPLINE), st.disc_spline)
        assert np.array_equal(b.derived(k, E.DERIVED_SUSCEPTIBILITY), st.susceptibility)
        assert np.array_equal(b.derived(k, E.DERIVED_AREAFACTOR), st.areafactor)
        assert np.array_equal(b.derived(k, E.DERIVED_BETA), st.beta)


@pytest.mark.parametrize("name", NAMES)
def test_deterministic_trajectory(name):
    """dynamics = deterministic (state.cc:627): no sampler involved.
    Engine vs oracle (same op order) and vs the reference golden trajectory."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    md = _variant(m, stochastic=0)
    eng = _engine(md)
    nth = d["theta"].shape[0]
    b = eng.batch(nth, E.KEEP_TRANSMEAN | E.KEEP_POP)
    b.set_params(d["theta"])
    b.simulate()
    Lh, tab = b.obs_loglik(with_tables=True)
    ov = b.obs_values()
    Li = b.latent_loglik()
    st = OracleState(md, stochastic=0, imap_from_pop=True)
    for k in range(nth):
        p = f"th{k}.det."
        tn = b.trajectory(k, E.TRAJ_TRANSNUM)
        ref_tn = d[p + "transnum"].reshape(m.traj_shape)
        scale = float(np.max(np.abs(ref_tn)))
        # free-running vs the reference: rounding differences grow like the epidemic (see test_oracle_golden)
        assert relerr(tn, ref_tn, floor=1e-6 * scale) < 1e-8
        st.set_param(d["theta"][k])
        st.simulate()
        assert relerr(tn, st.transnum, floor=1e-9 * scale) < RTOL
        assert relerr(b.trajectory(k, E.TRAJ_POP), st.pop, floor=1e-9 * scale) < RTOL
        assert relerr(ov[k], st.obs_values(), floor=1e-9) < RTOL
        assert relerr(Lh[k], st.obs_loglik()) < RTOL
        assert relerr(tab[k], st.EF_datatable(), floor=1e-9) < RTOL
        assert relerr(Li[k], st.latent_loglik()) < RTOL
        assert relerr(-2 * Lh[k], d[p + "EF"]) < 1e-7


@pytest.mark.parametrize("name", NAMES)
def test_teacher_forced_on_reference_trajectory(name):
    """One-step maps on the REFERENCE's own trajectory: import its event counts,
    replay (State::initialise_from_particle, state.cc:294-322) and compare the
    transition means, observation sums / likelihoods and the latent likelihood
    with what the reference itself computed -- 1e-10."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    eng = _engine(_variant(m, stochastic=1))
    nth = d["theta"].shape[0]
    b = eng.batch(nth, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(d["theta"])
    for k in range(nth):
        b.import_particle(k, d["theta"][k], d[f"th{k}.sto.transnum"].reshape(m.traj_shape).astype(np.float64))
    b.replay()
    Lh, tab = b.obs_loglik(with_tables=True)
    ov = b.obs_values()
    Li = b.latent_loglik()
    Lsec = np.stack([b.obs_loglik(section=s) for s in range(eng.nsection)], axis=1)
    assert eng.nsection == int(d["meta.nsection"][0])
    for k in range(nth):
        p = f"th{k}.sto."
        scale = float(np.max(np.abs(d[p + "transmean"])))
        assert relerr(b.trajectory(k, E.TRAJ_TRANSMEAN).ravel(), d[p + "transmean"], floor=1e-6 * scale) < RTOL
        assert relerr(ov[k], d[p + "obs_value"], floor=1e-9) < RTOL
        assert relerr(-2 * Lh[k], d[p + "EF"]) < RTOL
        assert relerr(tab[k], d[p + "EF_datatable"], floor=1e-9) < RTOL
        assert relerr(Li[k], d[p + "Li"]) < RTOL
        assert relerr(Lsec[k], d[p + "L_section"], floor=1e-9) < RTOL
        th, tn = b.export_particle(k)
        assert np.array_equal(th, d["theta"][k])
        assert np.array_equal(tn.ravel(), d[p + "transnum"].ravel().astype(np.float64))


@pytest.mark.parametrize("name", NAMES)
def test_stochastic_counts_bit_exact(name):
    """Integer event counts of beep_simulate == the oracle's, bit for bit, under
    the shared per-cell Philox streams; likelihoods on those counts to 1e-10."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    ms = _variant(m, stochastic=1)
    eng = _engine(ms)
    nth = d["theta"].shape[0]
    P = 2 * nth
    theta = np.concatenate([d["theta"], d["theta"]])
    ids = np.arange(P, dtype=np.uint64) * 1000003 + (1 << 40)
    b = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(theta)
    b.simulate(seed=2024, stream_ids=ids)
    Lh, tab = b.obs_loglik(with_tables=True)
    ov = b.obs_values()
    Li = b.latent_loglik()
    st = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    for k in range(P):
        st.set_param(theta[k])
        st.simulate(seed=2024, stream_id=int(ids[k]))
        tn = b.trajectory(k, E.TRAJ_TRANSNUM)
        assert np.array_equal(tn.astype(np.int64), st.transnum.astype(np.int64)), f"particle {k}"
        scale = float(np.max(np.abs(st.transmean)))
        assert relerr(b.trajectory(k, E.TRAJ_TRANSMEAN), st.transmean, floor=1e-9 * scale) < RTOL
        assert relerr(ov[k], st.obs_values(), floor=1e-9) < RTOL
        assert relerr(Lh[k], st.obs_loglik()) < RTOL
        assert relerr(tab[k], st.EF_datatable(), floor=1e-9) < RTOL
        assert relerr(Li[k], st.latent_loglik()) < RTOL
    # different streams give different trajectories; same stream reproduces
    assert not np.array_equal(b.trajectory(0), b.trajectory(nth))
    b2 = eng.batch(P, E.KEEP_TRANSNUM)
    b2.set_params(theta)
    b2.simulate(seed=2024, stream_ids=ids)
    assert np.array_equal(b2.trajectory(1), b.trajectory(1))


@pytest.mark.parametrize("name", NAMES)
def test_sectioned_simulation(name):
    """PMCMC::obs_prob section loop (pmcmc.cc:190-199): simulate(ti,tf) section by
    section gives the same trajectory as one sweep, and per-section likelihoods
    equal the oracle's calculate_section."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    ms = _variant(m, stochastic=1)
    eng = _engine(ms)
    P = 3
    theta = np.repeat(d["theta"][:1], P, axis=0)
    whole = eng.batch(P, E.KEEP_TRANSNUM)
    whole.set_params(theta)
    whole.simulate(seed=5)
    sec = eng.batch(P, E.KEEP_TRANSNUM)
    sec.set_params(theta)
    st = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    st.set_param(theta[0])
    st.simulate(seed=5, stream_id=1)
    Ls = []
    for s in range(eng.nsection):
        ti, tf = eng.section_range(s)
        assert (ti, tf) == st.section_range(s)
        sec.simulate(ti, tf, seed=5)
        Ls.append(sec.obs_loglik(section=s))
    Ls = np.stack(Ls, axis=1)
    for k in range(P):
        assert np.array_equal(sec.trajectory(k), whole.trajectory(k))
    assert np.array_equal(sec.trajectory(1).astype(np.int64), st.transnum.astype(np.int64))
    ref = [st.obs_loglik(s) for s in range(st.nsection)]
    assert relerr(Ls[1], ref, floor=1e-9) < RTOL


@pytest.mark.parametrize("sync", [None, 1, 2])
@pytest.mark.parametrize("name", NAMES)
def test_mbp(name, sync):
    """K3 vs the oracle's Mbp::mbp (mbp.cc:252-330): pure MBP, mixed SIMU/MBP,
    acceptance bookkeeping, and the fixed-event parameter proposal (mbp.cc:460-491) --
    in the flavour the launcher picks for the model and with the CTA- and grid-synchronised
    flavours forced (k_traj<MBP|REPLAY, SYNC, KEEP>: every instantiation the bench reports)."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    if d["theta"].shape[0] < 2:
        pytest.skip("single parameter vector")
    ms = _variant(m, stochastic=1)
    eng = _engine(ms, force_sync=sync)
    if sync == 2:
        # a demographic change (state.cc:339-376) keeps the particle where its areas can be summed each step, whatever is forced
        assert eng.sync_scope == ("GRID" if ms.arrays["dcc_info"].size == 0 else ("CTA" if ms.A > 1 else "NONE"))
    P = 4
    th0 = np.repeat(d["theta"][:1], P, axis=0)
    th1 = np.repeat(d["theta"][1:2], P, axis=0)
    ids = np.arange(P, dtype=np.uint64) + 77
    flags = E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN
    init = eng.batch(P, flags)
    prop = eng.batch(P, flags)
    init.set_params(th0)
    init.simulate(seed=1, stream_ids=ids)
    # oracle twin of particle 2
    k = 2
    oi = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    op = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    oi.set_param(th0[k])
    oi.simulate(seed=1, stream_id=int(ids[k]))
    assert np.array_equal(init.trajectory(k).astype(np.int64), oi.transnum.astype(np.int64))

    # (a) pure MBP towards theta1
    prop.set_params(th1)
    prop.mbp(init, None, seed=9, stream_ids=ids)
    op.set_param(th1[k])
    op.mbp(oi, None, seed=9, stream_id=int(ids[k]))
    assert np.array_equal(prop.trajectory(k).astype(np.int64), op.transnum.astype(np.int64))
    scale = float(np.max(np.abs(op.transmean)))
    assert relerr(prop.trajectory(k, E.TRAJ_TRANSMEAN), op.transmean, floor=1e-9 * scale) < RTOL
    Lp = prop.obs_loglik()
    assert relerr(Lp[k], op.obs_loglik()) < RTOL
    assert relerr(prop.latent_loglik()[k], op.latent_loglik()) < RTOL

    # (b) mixed SIMU / MBP, unchanged parameters: MBP cells reproduce the initial counts
    T, A = m.T, m.A
    mask = np.zeros((T, A), dtype=np.uint8)
    mask[T // 3:(2 * T) // 3, ::2] = 1
    prop.set_params(th0)
    prop.mbp(init, mask, seed=10, stream_ids=ids)
    op.set_param(th0[k])
    op.mbp(oi, mask, seed=10, stream_id=int(ids[k]))
    tp = prop.trajectory(k)
    assert np.array_equal(tp.astype(np.int64), op.transnum.astype(np.int64))
    assert np.array_equal(tp[:T // 3], init.trajectory(k)[:T // 3])

    # accept particles 0 and 2: their slots of `init` take the proposal, the others keep theirs
    before1 = init.trajectory(1)
    acc = np.array([1, 0, 1, 0], dtype=np.uint8)
    init.accept_from(prop, acc)
    assert np.array_equal(init.trajectory(k), tp)
    assert np.array_equal(init.trajectory(1), before1)
    assert relerr(init.obs_loglik()[k], op.obs_loglik()) < RTOL

    # (c) parameter proposal with the events fixed
    prop.set_params(th1)
    for s in range(P):
        prop.import_particle(s, th1[s], init.trajectory(s))
    prop.replay()
    o3 = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    o3.set_param(th1[k])
    o3.init_from_particle(init.trajectory(k))
    assert relerr(prop.trajectory(k, E.TRAJ_TRANSMEAN), o3.transmean, floor=1e-9 * scale) < RTOL
    assert relerr(prop.latent_loglik()[k], o3.latent_loglik()) < RTOL


@pytest.mark.parametrize("name", ["EX_C2", "EX_C5", "EX_C4", "EX_C6"])
@pytest.mark.parametrize("sync", [1, 2])
def test_sync_flavours_agree(name, sync):
    """The CTA-synchronised and grid-synchronised kernels give the same bits as
    the default choice (and hence as the oracle)."""
    from beep_b200 import engine as E
    m, d = load_golden(name)
    ms = _variant(m, stochastic=1)
    P = 5
    theta = np.repeat(d["theta"][:1], P, axis=0)
    e0 = _engine(ms)
    b0 = e0.batch(P, E.KEEP_TRANSNUM)
    b0.set_params(theta)
    b0.simulate(seed=3)
    e1 = _engine(ms, force_sync=sync)
    b1 = e1.batch(P, E.KEEP_TRANSNUM)
    b1.set_params(theta)
    b1.simulate(seed=3)
    for k in range(P):
        assert np.array_equal(b0.trajectory(k), b1.trajectory(k))
    assert relerr(b0.obs_loglik(), b1.obs_loglik()) < 1e-12
    assert np.array_equal(b0.pop(2), b1.pop(2))


def test_permute_and_section_state():
    """Mpi::copy_particles (mpi.cc:81-165) and end_sec_swap (mpi.cc:215-327) on one GPU."""
    from beep_b200 import engine as E
    m, d = load_golden("EX_C4")
    ms = _variant(m, stochastic=1)
    eng = _engine(ms)
    P = 6
    theta = np.repeat(d["theta"][:1], P, axis=0)
    b = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(theta)
    ti, tf = eng.section_range(0)
    b.simulate(ti, tf, seed=11)
    pops = [b.pop(k) for k in range(P)]
    back = np.array([0, 0, 2, 5, 5, 5], dtype=np.uint32)
    b.permute(back, E.PERMUTE_SECTION_STATE)
    for k in range(P):
        assert np.array_equal(b.pop(k), pops[back[k]])
    # continue: particles 3,4,5 now share an ancestor state but draw from their own streams
    if eng.nsection > 1:
        t2 = eng.section_range(1)
        b.simulate(t2[0], t2[1], seed=11)
        assert not np.array_equal(b.trajectory(3)[t2[0]:t2[1]], b.trajectory(4)[t2[0]:t2[1]])
    # full copy
    full = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    full.set_params(theta)
    full.simulate(seed=12)
    trajs = [full.trajectory(k) for k in range(P)]
    L0 = full.obs_loglik()
    src = np.array([E.KEEP, 0, 1, 1, E.KEEP, 4], dtype=np.uint32)
    full.permute(src, E.PERMUTE_FULL)
    for k in range(P):
        s = k if src[k] == E.KEEP else src[k]
        assert np.array_equal(full.trajectory(k), trajs[s])
    L1 = full.obs_loglik()
    for k in range(P):
        s = k if src[k] == E.KEEP else src[k]
        assert L1[k] == L0[s]


@pytest.mark.parametrize("name", ["EX_A1", "EX_C4", "EX_C2"])
def test_particle_filter_matches_oracle(name):
    """PMCMC::obs_prob (pmcmc.cc:180-268) through beep_b200.shard.ShardedFilter: the engine and the
    oracle give the same ancestry (bit-exact) and the same log marginal likelihood (1e-10)."""
    from beep_b200 import engine as E
    from beep_b200.shard import ShardedFilter
    from oracle_batch import OracleBatch, OracleEngine
    m, d = load_golden(name)
    ms = _variant(m, stochastic=1)
    N = 6
    fe = ShardedFilter(_engine(ms).batch(N))
    fo = ShardedFilter(OracleBatch(OracleEngine(ms), N))
    le = fe.run(d["theta"][0], seed=4)
    lo = fo.run(d["theta"][0], seed=4)
    assert [a.tolist() for a in fe.ancestry] == [a.tolist() for a in fo.ancestry]
    assert relerr(le, lo) < RTOL
    # section-state exchange through host buffers (the multi-rank path) == device permute
    b = fe.batch
    pops = b.pop_all()
    b.set_pop_all(pops[::-1].copy(), b.model.T)
    assert np.array_equal(b.pop(0), pops[N - 1])


def test_cpp_host_filter(tmp_path):
    """The C++ host layer (include/beep_b200.hpp, host/pmcmc_filter.cc) drives the same C ABI and
    reproduces the Python-driven particle filter to the last bit of the ancestry (same log-likelihood)."""
    import subprocess
    from beep_b200 import engine as E
    from beep_b200.bpk import write_bpk
    from beep_b200.shard import ShardedFilter
    from conftest import ROOT
    exe = os.path.join(ROOT, "host", "pmcmc_filter")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    m, d = load_golden("EX_C4")
    ms = _variant(m, stochastic=1)
    arrays = ms.to_prefixed()
    arrays["theta"] = np.ascontiguousarray(d["theta"][0], dtype=np.float64)
    path = str(tmp_path / "model.bpk")
    write_bpk(path, arrays)
    P = 64
    out = subprocess.run([exe, path, str(P), "1", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    ll_cpp = float(out.stdout.split("log-likelihood=")[1].split()[0])
    ll_py = ShardedFilter(_engine(ms).batch(P)).run(d["theta"][0], seed=4)
    assert relerr(ll_cpp, ll_py) < 1e-12


def test_many_particles():
    """More particles than a grid's y-dimension allows (65535): the particle index rides on grid.x."""
    from beep_b200 import engine as E
    m, d = load_golden("EX_A1")
    ms = _variant(m, stochastic=1)
    eng = _engine(ms)
    P = 70000
    b = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(d["theta"][0])
    b.simulate(seed=17)
    L = b.obs_loglik()
    Li = b.latent_loglik()
    st = OracleState(ms, stochastic=1, keyed=True, imap_from_pop=True)
    st.set_param(d["theta"][0])
    for k in (0, 65535, 65536, P - 1):
        st.simulate(seed=17, stream_id=k)
        assert np.array_equal(b.trajectory(k).astype(np.int64), st.transnum.astype(np.int64))
        assert relerr(L[k], st.obs_loglik()) < RTOL
        assert relerr(Li[k], st.latent_loglik()) < RTOL
    src = np.full(P, E.KEEP, dtype=np.uint32)
    src[P - 1] = 3
    t3 = b.trajectory(3)
    b.permute(src, E.PERMUTE_FULL)
    assert np.array_equal(b.trajectory(P - 1), t3)


def test_error_paths():
    from beep_b200 import engine as E
    m, d = load_golden("EX_sim")
    eng = _engine(m)
    b = eng.batch(2)
    with pytest.raises(E.BeepError) as e:
        b.simulate()
    assert e.value.code == -5            # BEEP_ERR_STATE: simulate before set_params
    b.set_params(d["theta"][0])
    with pytest.raises(E.BeepError):
        b.simulate(5, 3)
    with pytest.raises(E.BeepError) as e:
        b.latent_loglik()
    assert e.value.code == -5
    bad = _variant(m, n_unsupported=1)
    with pytest.raises(E.BeepError) as e:
        E.Engine(bad)
    assert e.value.code == -2


def test_stochastic_moments_vs_reference_sampler():
    """Monte-Carlo agreement with the reference's own std:: samplers: final sizes of
    EX.toml over many replicates (engine) vs the reference-stream golden value is
    covered in tests/test_statistical_gpu.py; here: Poisson mean/variance of one step."""
    from beep_b200 import engine as E
    m, d = load_golden("EX_sim")
    ms = _variant(m, stochastic=1)
    eng = _engine(ms)
    P = 4096
    b = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(d["theta"][0])
    b.simulate(0, 1, seed=99)
    tn = np.stack([b.trajectory(k)[0].ravel() for k in range(0, P, 16)])
    tm = b.trajectory(0, E.TRAJ_TRANSMEAN)[0].ravel()
    nz = tm > 0
    mean = tn[:, nz].mean(axis=0)
    assert np.all(np.abs(mean - tm[nz]) < 6 * np.sqrt(tm[nz] / tn.shape[0]) + 1e-12)
