"""mode=prediction on the engine: the model modifications of model.modelmod (reference src/model.cc:1044-1050, 1349-1357,
1440-1585; state.cc:415, 446-453) and the restart in mid-trajectory of Simulate::model_modification (simulate.cc:141-165:
State::load of a posterior sample, then State::simulate(pred_start, ndivision)).

The fixtures PRED_F3 / PRED_C2 come from the reference constructed in prediction mode on TOML files that carry every kind of
modification (tests/golden/make_golden.py); tests/test_oracle_golden.py pins the oracle to the reference's own set_param,
simulate and Mbp::mbp under them, tests/test_engine_gpu.py (which runs over every fixture) the engine to the oracle.  Here:
the derived tables against the reference directly, and the prediction restart against the oracle."""
import numpy as np
import pytest

from beep_b200 import engine as E
from conftest import load_golden, relerr
from oracle_lib import OracleState

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["PRED_F3", "PRED_C2"])
def test_modified_tables_match_reference(name):
    m, d = load_golden(name)
    assert m.mod_transmean.size and m.mod_efoi.size
    eng = E.Engine(m, 0)
    assert eng.chunk_tiles == 0                      # models with modifications run the general tile kernel
    K = d["theta"].shape[0]
    bt = eng.batch(K)
    bt.set_params(d["theta"])
    for k in range(K):
        assert relerr(bt.derived(k, E.DERIVED_DISC_SPLINE).ravel(), d[f"th{k}.disc_spline"]) < 1e-10
        assert relerr(bt.derived(k, E.DERIVED_BETA).ravel(), d[f"th{k}.beta"]) < 1e-10


@pytest.mark.parametrize("name", ["PRED_F3", "PRED_C2"])
def test_prediction_restart_matches_oracle(name):
    """posterior sample (here: the reference's own stochastic trajectory of the fixture) -> replay up to the start of the
    prediction -> simulate the rest under the modifications: counts bit-exact against the oracle doing the same"""
    m, d = load_golden(name)
    eng = E.Engine(m, 0)
    T = m.T
    ts = (2 * T) // 3
    tn_post = d["th0.sto.transnum"].reshape(m.traj_shape).astype(np.float64).copy()
    tn_post[ts:] = 0                                  # State::load pads a shorter posterior with zeros (state.cc:794-805)
    P = 4
    bt = eng.batch(P, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    for k in range(P):
        bt.import_particle(k, d["theta"][0], tn_post)
    bt.replay_range(0, ts)
    bt.simulate(ts, T, seed=77)
    L = bt.obs_loglik()
    for k in (0, P - 1):
        o = OracleState(m, stochastic=1, keyed=True, imap_from_pop=True)
        o.set_param(d["theta"][0])
        o.init_from_particle(tn_post)
        o.simulate(ts, T, seed=77, stream_id=k)
        tn = bt.trajectory(k).astype(np.int64)
        assert np.array_equal(tn[:ts], tn_post[:ts].astype(np.int64))
        assert np.array_equal(tn, o.transnum.astype(np.int64)), "prediction counts differ from the oracle"
        assert tn[ts:].sum() > 0
        assert abs(L[k] - o.obs_loglik()) <= 1e-10 * abs(o.obs_loglik())
