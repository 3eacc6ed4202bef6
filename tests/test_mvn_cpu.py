"""ABC-SMC weight kernel -- the CPU oracle against the reference.

tests/golden/mvn_*.npz hold the reference's OWN MVN::setup / propose / probability outputs and the
sums of ABCSMC::calculate_particle_weight (abcsmc.cc:264-268), produced by oracle/ref/ref_mvn.cc
linked against the unmodified reference objects (tests/golden/make_golden.py).  The oracle restates
mvn.cc:185-196 in the reference's loop order."""
import numpy as np
import pytest

from conftest import ROOT
import oracle_lib

CASES = ["mvn_small", "mvn_england"]


def load(name):
    import os
    return dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    d = load(name)
    s = oracle_lib.mvn_kernel_sums(d["var"], d["inv_M"], d["size"][0], d["param_prop"], d["param_last"], d["w_last"])
    # same operations in the same order; the reference is built with -mfma and may contract: allow a few ulps of the exponent
    assert np.max(np.abs(s - d["sums"]) / d["sums"]) < 1e-12


def test_golden_is_a_hard_case():
    """The England_AS-like fixture is ill conditioned (cond(M) ~ 5e5) with sums down at 1e-30: the regime
    in which a whitened evaluation has to prove itself."""
    d = load("mvn_england")
    assert np.linalg.cond(d["M"]) > 1e5
    assert d["sums"].max() < 1e-20 and d["sums"].min() > 0
    assert np.max(np.abs(d["inv_M"] @ d["M"] - np.eye(d["M"].shape[0]))) < 1e-8
