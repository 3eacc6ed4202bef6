"""Pins the CPU oracle (oracle/beep_oracle.c) against the UNMODIFIED reference.

The fixtures in tests/golden/ were produced by the reference's own
State::set_param / State::simulate / ObservationModel / Mbp::mbp (see
tests/golden/make_golden.py and oracle/ref/ref_harness.cc).  Tolerances:
  - doubles: 1e-10 relative (BASELINE.json north_star), with an absolute floor
    for quantities that pass through zero;
  - integer event counts under the shared injected stream: bit-exact.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, relerr
from oracle_lib import OracleState

RTOL = 1e-10
NAMES = golden_names()


def ntheta(d):
    return d["theta"].shape[0]


@pytest.mark.parametrize("name", NAMES)
def test_set_param(name):
    """State::set_param outputs, reference src/state.cc:66-91"""
    m, d = load_golden(name)
    st = OracleState(m)
    for k in range(ntheta(d)):
        st.set_param(d["theta"][k])
        p = f"th{k}."
        # the reference leaves the infection transition's row unsized (model.cc:1127-1128)
        assert relerr(st.transrate[m.trans_inf == 0].ravel(), d[p + "transrate"]) < RTOL
        assert relerr(st.disc_spline.ravel(), d[p + "disc_spline"]) < RTOL
        assert relerr(st.susceptibility, d[p + "susceptibility"]) < RTOL
        assert relerr(st.areafactor.ravel(), d[p + "areafactor"]) < RTOL
        assert relerr(st.beta.ravel(), d[p + "beta"]) < RTOL
        T = m.T
        nt3 = np.concatenate([st.Ntime[s].ravel() for s in (0, T // 2, T - 1)])
        assert relerr(nt3, d[p + "Ntime3"]) < RTOL


@pytest.mark.parametrize("name", NAMES)
def test_deterministic_trajectory(name):
    """dynamics = deterministic: transnum = transmean (state.cc:627), no sampler.

    (1) Teacher-forced on the reference's own trajectory (its pop, Imap, Idiag):
        the one-step maps -- force of infection, transition means, observation
        sums, observation log-likelihood, latent likelihood -- match to 1e-10.
    (2) Free-running over all T steps: an epidemic is an exponentially growing
        recurrence, so last-ulp differences (the reference is built -O3 -mfma with
        FMA contraction, CMakeLists.txt:27) grow ~e^(rt); 1e-8 is the bound there.
    """
    m, d = load_golden(name)
    st = OracleState(m, stochastic=0)
    for k in range(ntheta(d)):
        p = f"th{k}.det."
        ref_tn = d[p + "transnum"].reshape(m.traj_shape)
        scale = float(np.max(np.abs(ref_tn)))
        st.set_param(d["theta"][k])
        st.simulate()
        assert relerr(st.transnum, ref_tn, floor=1e-6 * scale) < 1e-8
        if p + "pop" not in d:
            continue
        st.pop[:] = d[p + "pop"].reshape(m.pop_shape)
        st.Imap[:] = d[p + "Imap"].reshape(st.Imap.shape)
        st.Idiag[:] = d[p + "Idiag"].reshape(st.Idiag.shape)
        st.transnum[:] = ref_tn
        st.recompute_transmean()
        assert relerr(st.transmean, ref_tn, floor=1e-9 * scale) < RTOL
        assert relerr(st.obs_values(), d[p + "obs_value"], floor=1e-9) < RTOL
        assert relerr(-2 * st.obs_loglik(), d[p + "EF"]) < RTOL
        assert relerr(st.EF_datatable(), d[p + "EF_datatable"], floor=1e-9) < RTOL
        assert relerr(st.latent_loglik(), d[p + "Li"]) < RTOL
        Ls = [st.obs_loglik(s) for s in range(st.nsection)]
        assert relerr(Ls, d[p + "L_section"], floor=1e-9) < RTOL


@pytest.mark.parametrize("name", NAMES)
def test_stochastic_counts_bit_exact(name):
    """Integer event counts under the shared injected sequential stream are
    bit-exact against the reference's State::simulate (state.cc:596-638)."""
    m, d = load_golden(name)
    st = OracleState(m, stochastic=1, keyed=False, imap_from_pop=False)
    seed = int(d["meta.seed"][0])
    for k in range(ntheta(d)):
        p = f"th{k}.sto."
        st.set_param(d["theta"][k])
        n0 = st.ndraw
        st.simulate(seed=seed, stream_id=k)
        assert st.ndraw - n0 == int(d[p + "ndraw"][0])
        assert np.array_equal(st.transnum.ravel().astype(np.int64), d[p + "transnum"].ravel().astype(np.int64))
        scale = float(np.max(np.abs(d[p + "transmean"])))
        assert relerr(st.transmean.ravel(), d[p + "transmean"], floor=1e-6 * scale) < RTOL
        assert relerr(st.obs_values(), d[p + "obs_value"], floor=1e-9) < RTOL
        assert relerr(-2 * st.obs_loglik(), d[p + "EF"]) < RTOL
        assert relerr(st.latent_loglik(), d[p + "Li"]) < RTOL
        Ls = [st.obs_loglik(s) for s in range(st.nsection)]
        assert relerr(Ls, d[p + "L_section"], floor=1e-9) < RTOL


@pytest.mark.parametrize("name", NAMES)
def test_sectioned_simulation(name):
    """PMCMC::obs_prob section loop (pmcmc.cc:190-199): simulate(ti,tf) per section"""
    m, d = load_golden(name)
    st = OracleState(m, stochastic=1)
    seed = int(d["meta.seed"][0])
    assert st.nsection == int(d["meta.nsection"][0])
    st.set_param(d["theta"][0])
    Ls = []
    for s in range(st.nsection):
        ti, tf = st.section_range(s)
        assert (ti, tf) == (int(d["meta.section_ti"][s]), int(d["meta.section_tf"][s]))
        st.simulate(ti, tf, seed=seed, stream_id=1000)
        Ls.append(st.obs_loglik(s))
    # the harness keys ONE stream for the whole section loop; the oracle re-keys per
    # call, so only the first section shares draws -- compare it bit-exactly
    ti, tf = st.section_range(0)
    ref = d["th0.sec.transnum"].reshape(m.traj_shape)
    assert np.array_equal(st.transnum[ti:tf].astype(np.int64), ref[ti:tf].astype(np.int64))
    assert relerr(Ls[0], d["th0.sec.L_section"][0]) < RTOL


@pytest.mark.parametrize("name", [n for n in NAMES if "mbp.a.transnum" in load_golden(n)[1]])
def test_mbp(name):
    """Mbp::mbp (mbp.cc:252-330): pure MBP, mixed SIMU/MBP, and the fixed-event
    parameter proposal (mbp.cc:460-491), counts bit-exact under the shared stream."""
    m, d = load_golden(name)
    seed = int(d["meta.seed"][0])
    init = OracleState(m, stochastic=1)
    prop = OracleState(m, stochastic=1)
    init.set_param(d["theta"][0])
    init.init_from_particle(d["th0.sto.transnum"].reshape(m.traj_shape).astype(np.float64))
    scale = float(np.max(np.abs(d["mbp.init.transmean"])))
    assert relerr(init.transmean.ravel(), d["mbp.init.transmean"], floor=1e-6 * scale) < RTOL
    assert relerr(init.latent_loglik(), d["mbp.init.Li"]) < RTOL

    # (a) pure MBP towards theta1
    prop.set_param(d["theta"][1])
    n0 = prop.ndraw
    prop.mbp(init, None, seed=seed, stream_id=2000)
    assert prop.ndraw - n0 == int(d["mbp.a.ndraw"][0])
    assert np.array_equal(prop.transnum.ravel().astype(np.int64), d["mbp.a.transnum"].ravel().astype(np.int64))
    # the reference forms Imap_p = Imap_i + dImap (state.cc:109-123), a cancelling sum whose
    # rounding noise is ~1e-16 |Imap_i|: measure against max(|x|, 1e-4 max|x|)
    assert relerr(prop.transmean.ravel(), d["mbp.a.transmean"], floor=1e-4 * scale) < RTOL
    assert relerr(-2 * prop.obs_loglik(), d["mbp.a.EF"]) < RTOL
    assert relerr(prop.latent_loglik(), d["mbp.a.Li"]) < RTOL

    # (b) mixed SIMU / MBP with unchanged parameters
    prop.set_param(d["theta"][0])
    prop.mbp(init, d["mbp.b.mask"], seed=seed, stream_id=3000)
    assert np.array_equal(prop.transnum.ravel().astype(np.int64), d["mbp.b.transnum"].ravel().astype(np.int64))
    assert relerr(-2 * prop.obs_loglik(), d["mbp.b.EF"]) < RTOL

    # (c) parameter proposal with the events fixed
    prop.copy_from(init)
    prop.set_param(d["theta"][1])
    prop.recompute_transmean()
    assert relerr(prop.transmean.ravel(), d["mbp.c.transmean"], floor=1e-6 * scale) < RTOL
    assert relerr(prop.latent_loglik(), d["mbp.c.Li"]) < RTOL


def test_imap_from_pop_equals_incremental():
    """The engine recomputes the infectivity map from the population every step;
    the reference updates it incrementally (state.cc:164-219).  Both agree to
    rounding (the reference's own check allows 1e-5, state_check.cc)."""
    for name in ("EX_C2", "EX_C4", "EX_C5"):
        m, d = load_golden(name)
        a = OracleState(m, stochastic=0, imap_from_pop=False)
        b = OracleState(m, stochastic=0, imap_from_pop=True)
        a.set_param(d["theta"][0]); b.set_param(d["theta"][0])
        a.simulate(); b.simulate()
        scale = float(np.max(np.abs(a.transnum)))
        assert relerr(a.transnum.ravel(), b.transnum.ravel(), floor=1e-6 * scale) < 1e-9
