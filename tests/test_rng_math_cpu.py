"""The software exp / log of include/beep_rng.h (shared by the CUDA kernels, the CPU oracle and
the reference harness, so parity tests alone could not see an inaccurate one): pinned against
libm to 1 ulp-level accuracy over the ranges the samplers and create_areafactor use."""
import numpy as np

from oracle_lib import lib


def _ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_beep_exp_accuracy():
    rng = np.random.default_rng(1)
    x = np.concatenate([-rng.uniform(0, 10, 200000), rng.uniform(-700, 700, 200000), rng.uniform(-1e-3, 1e-3, 50000),
                        np.array([0.0, -0.0, 1.0, -1.0, 700.0, -700.0, np.log(2) / 2, -np.log(2) / 2])])
    out = np.empty_like(x)
    lib().oracle_exp_many(x.ctypes.data, x.size, out.ctypes.data)
    u = _ulps(out, np.exp(x))
    assert u.max() <= 2.0, (u.max(), x[u.argmax()])
    assert np.mean(u <= 1.0) > 0.99


def test_beep_log_accuracy():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(1e-300, 1, 100000), rng.uniform(1, 1e6, 100000), 10.0 ** rng.uniform(-300, 300, 100000),
                        np.array([1.0, 2.0, 0.5, 5e-324, 1e-310])])
    out = np.empty_like(x)
    lib().oracle_log_many(x.ctypes.data, x.size, out.ctypes.data)
    ref = np.log(x)
    ok = ref != 0
    u = _ulps(out[ok], ref[ok])
    assert u.max() <= 2.0, (u.max(), x[ok][u.argmax()])
    assert np.all(out[~ok] == 0)
