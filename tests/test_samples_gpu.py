"""Posterior sample files: State::load / State::save (state.cc:709-790) through the
engine.  tests/golden/samples_EX_A1/sample7_* were written by the reference itself (the mode=mcmcmbp run of
tests/golden/make_golden.py); the engine loads them into a slot, rebuilds the state by replay
(State::initialise_from_particle) and writes them back: the files must be identical byte for byte, and the
error function recomputed from the loaded events must be the stored one (to the 6 digits the file keeps)."""
import filecmp
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
NAMES = ["zero", "I mean occupancy", "I Infectivity", "R-spline", "one", "efoi-spline"]      # model.param[].name of EX_A1.toml


def test_sample_file_round_trip(tmp_path):
    from beep_b200 import engine as E
    from beep_b200.model import ModelDesc
    m, d = load_golden("EX_A1")
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    eng = E.Engine(ModelDesc(arrays))
    b = eng.batch(3, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    b.set_params(d["theta"][0])
    b.simulate(seed=1)
    src = os.path.join(ROOT, "tests", "golden", "samples_EX_A1", "sample7")
    EF = b.load_sample(1, src, NAMES)
    assert EF == 308.913
    b.replay()
    L = b.obs_loglik()
    assert abs(-2 * L[1] - EF) < 2e-3 * EF           # parameters and EF are stored with 6 significant digits
    # the derived tables follow the loaded parameters (State::initialise_from_particle starts with set_param,
    # state.cc:296): the latent-process likelihood of the loaded events equals the oracle's
    from oracle_lib import OracleState
    th1, tn1 = b.export_particle(1)
    st = OracleState(ModelDesc(arrays), stochastic=1, keyed=True, imap_from_pop=True)
    st.set_param(th1)
    st.init_from_particle(tn1)
    Li = b.latent_loglik()
    assert abs(Li[1] - st.latent_loglik()) <= 1e-10 * abs(st.latent_loglik())
    assert abs(L[1] - st.obs_loglik()) <= 1e-10 * abs(st.obs_loglik())
    out = str(tmp_path / "sample7")
    b.save_sample(1, out, NAMES, EF, start_time=0.0)
    for suffix in ("_parameter.csv", "_transition.csv"):
        assert filecmp.cmp(src + suffix, out + suffix, shallow=False), suffix
    # the neighbouring slots were not touched
    th0, tn0 = b.export_particle(0)
    assert np.array_equal(th0, d["theta"][0])
    with pytest.raises(E.BeepError):
        b.load_sample(1, src, ["wrong"] + NAMES[1:])
