"""ABC-SMC weight kernel on the GPU (beep_mvn_kernel_sums, K7) against the reference's golden sums
and the CPU oracle.  Tolerance: 1e-10 relative (BASELINE.json north_star, deterministic FP64 pieces);
the engine evaluates the quadratic form in whitened coordinates, the reference as a double loop."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden
import oracle_lib

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _engine():
    from beep_b200 import engine as E
    m, _ = load_golden("EX_A1")
    return E, E.Engine(m)


@pytest.mark.parametrize("name", ["mvn_small", "mvn_england"])
def test_matches_reference_golden(name):
    E, eng = _engine()
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    s = E.mvn_kernel_sums(eng, d["var"], d["inv_M"], d["size"][0], d["param_prop"], d["param_last"], d["w_last"])
    assert np.max(np.abs(s - d["sums"]) / d["sums"]) < RTOL


def _random_case(rng, nparam, nvar, Nl, N, size=1.1):
    var = np.sort(rng.choice(nparam, nvar, replace=False)).astype(np.int32)
    load = rng.normal(size=(nparam, 5))
    base = rng.uniform(0.5, 4, nparam)
    last = base + 0.2 * (rng.normal(size=(Nl, 5)) @ load.T + rng.normal(size=(Nl, nparam)))
    w = rng.uniform(0.1, 1, Nl)
    mu = (w[:, None] * last[:, var]).sum(0) / w.sum()
    c = last[:, var] - mu
    M = (w[:, None] * c).T @ c / w.sum()
    if Nl < nvar + 2:                                  # too few samples for a covariance: any SPD matrix will do
        q = rng.normal(size=(nvar, nvar))
        M = q @ q.T / nvar + 0.05 * np.eye(nvar)
    inv_M = np.linalg.inv(M)
    chol = np.linalg.cholesky(M)
    pick = rng.integers(0, Nl, N)
    prop = last[pick].copy()
    prop[:, var] += size * (rng.normal(size=(N, nvar)) @ chol.T)
    return var, inv_M, size, prop, last, w


def test_matches_oracle_ragged_sizes():
    """Sizes that are not multiples of the 64 x 64 tile or of the 16-wide k chunk, one proposal, one ancestor."""
    E, eng = _engine()
    rng = np.random.default_rng(5)
    for nparam, nvar, Nl, N in [(9, 3, 1, 1), (30, 17, 65, 3), (50, 33, 257, 130), (40, 16, 64, 64)]:
        case = _random_case(rng, nparam, nvar, Nl, N)
        ref = oracle_lib.mvn_kernel_sums(*case[:2], case[2], case[3], case[4], case[5])
        got = E.mvn_kernel_sums(eng, *case)
        assert np.max(np.abs(got - ref) / ref) < RTOL


def test_generation_size_deterministic():
    """A generation's worth (4096 proposals x 2048 ancestors x 170 free parameters): the same bits twice,
    and a sample of rows against the oracle."""
    E, eng = _engine()
    rng = np.random.default_rng(6)
    case = _random_case(rng, 180, 170, 2048, 4096)
    a = E.mvn_kernel_sums(eng, *case)
    b = E.mvn_kernel_sums(eng, *case)
    assert np.array_equal(a, b)
    rows = [0, 63, 64, 1000, 4095]
    var, inv_M, size, prop, last, w = case
    ref = oracle_lib.mvn_kernel_sums(var, inv_M, size, prop[rows], last, w)
    assert np.max(np.abs(a[rows] - ref) / ref) < RTOL


def test_errors():
    E, eng = _engine()
    rng = np.random.default_rng(7)
    var, inv_M, size, prop, last, w = _random_case(rng, 12, 5, 20, 4)
    with pytest.raises(E.BeepError):
        E.mvn_kernel_sums(eng, var, -inv_M, size, prop, last, w)         # not positive definite
    bad = var.copy(); bad[0] = 99
    with pytest.raises(E.BeepError):
        E.mvn_kernel_sums(eng, bad, inv_M, size, prop, last, w)
