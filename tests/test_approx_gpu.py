"""Row N3 on the device: beep_likelihood_approx (beep_b200/csrc/approx.cuh) through the C ABI against the CPU oracle
(same sums in the same order: 1e-10 relative) and against the reference's own State::likelihood_approx
(tests/golden/APX_*.npz: 1e-9, see tests/test_approx_cpu.py for the tolerance)."""
import os

import numpy as np
import pytest

from beep_b200 import engine as E
from beep_b200.model import ModelDesc
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["APX_A1", "APX_B4", "APX_C4", "APX_C2", "APX_D3", "APX_C5"]
ILL = {("APX_C5", 1): 1e-6}


def load(name):
    d = dict(np.load(os.path.join(GOLD, name + ".npz")))
    return d, ModelDesc.from_prefixed(d)


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_and_oracle(name):
    d, m = load(name)
    eng = E.Engine(m, 0)
    assert eng.has_approx
    th = d["theta"]
    bt = eng.batch(th.shape[0])
    bt.set_params(th)
    invT = float(d["approx.invT"][0])
    L = bt.likelihood_approx(invT)
    o = OracleState(m)
    for k in range(th.shape[0]):
        Lr = float(d["approx.L"][k])
        Lo = o.likelihood_approx(th[k], invT)
        tol = ILL.get((name, k), 1e-9)
        assert abs(L[k] - Lr) <= tol * abs(Lr), (name, k, L[k], Lr)
        assert abs(L[k] - Lo) <= max(tol * 0.1, 1e-10) * abs(Lo), (name, k, L[k], Lo)


def test_batch_of_perturbed_parameters():
    """A CMA-ES generation: 96 parameter vectors around the model's own values, every L equal to the oracle's."""
    d, m = load("APX_C4")
    eng = E.Engine(m, 0)
    rng = np.random.default_rng(5)
    P = 96
    th = np.repeat(d["theta"][:1], P, axis=0)
    free = np.abs(d["theta"][1] - d["theta"][0]) > 0          # the parameters the prior varies
    th[:, free] *= 1 + 0.05 * rng.standard_normal((P, int(free.sum())))
    bt = eng.batch(P)
    bt.set_params(th)
    L = bt.likelihood_approx(1.0)
    assert np.all(np.isfinite(L)) and np.unique(L).size > P // 2
    o = OracleState(m)
    for k in (0, 17, 95):
        Lo = o.likelihood_approx(th[k], 1.0)
        assert abs(L[k] - Lo) <= 1e-10 * abs(Lo)
    # the same batch again: same bits (no atomics, fixed order of sums)
    assert np.array_equal(L, bt.likelihood_approx(1.0))


def test_unsupported_models_are_refused():
    d = dict(np.load(os.path.join(GOLD, "EX_C6.npz")))          # two strains
    m = ModelDesc.from_prefixed(d)
    eng = E.Engine(m, 0)
    assert not eng.has_approx
    bt = eng.batch(1)
    bt.set_params(d["theta"][:1])
    with pytest.raises(E.BeepError):
        bt.likelihood_approx(1.0)
