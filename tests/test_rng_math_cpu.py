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


def _presearch(u, lam, rel_err):
    """The device's single-precision pre-search of beep_poisson_inv_u (include/beep_rng.h), restated with numpy float32;
    rel_err perturbs exp(-lam) and every reciprocal by the worst relative error the hardware approximations may have.
    Returns (k, safe): where safe, the device returns k without running the exact search."""
    f = np.float32
    uf, lf = u.astype(f), lam.astype(f)
    p = (np.exp(-lf.astype(np.float64)) * (1 + rel_err)).astype(f)
    cdf, prev, kf = p.copy(), np.full(u.shape, -1, f), np.zeros(u.shape, f)
    while True:
        go = (uf > cdf) & (kf < 60)
        if not go.any(): break
        kf = np.where(go, kf + f(1), kf)
        rk = (1.0 / np.maximum(kf, 1).astype(np.float64) * (1 + 1.2e-7 * np.sign(rel_err))).astype(f)
        pn = (p * lf).astype(f) * rk
        p = np.where(go, pn.astype(f), p)
        prev = np.where(go, cdf, prev)
        cdf = np.where(go, (cdf + p).astype(f), cdf)
    tol = (f(5e-6) + f(1e-6) * kf).astype(f)
    safe = ((cdf - uf).astype(f) > tol) & ((uf - prev).astype(f) > tol)
    return kf.astype(np.int32), safe


def test_device_presearch_never_changes_a_draw():
    """Wherever the single-precision pre-search settles a draw, the exact search (the oracle's) returns the same k --
    on random uniforms, on uniforms placed at and around every step of the distribution function, and with the
    hardware approximations pushed to their error bounds in either direction; it settles all but ~1e-4 of the draws."""
    from scipy.stats import poisson
    rng = np.random.default_rng(7)
    n = 400000
    lam = np.concatenate([rng.uniform(0, 10, n), 10.0 ** rng.uniform(-8, 1, n)])
    lam = np.minimum(lam, np.nextafter(10.0, 0))
    u = (rng.integers(0, 1 << 52, lam.size).astype(np.float64) + 0.5) * 2.0 ** -52
    # adversarial: uniforms within a few 1e-5 / 1e-7 / 1e-12 of a step of the distribution function, and near 0 and 1
    la = np.repeat(rng.uniform(1e-3, 10, 4000), 50)
    ka = rng.integers(0, 40, la.size)
    ua = poisson.cdf(ka, la) + rng.choice([1e-4, 3e-5, 1e-5, 3e-6, 1e-7, 1e-12, 0.0], la.size) * rng.choice([-1, 1], la.size) * rng.uniform(0, 1, la.size)
    ua = np.clip(ua, 2.0 ** -53, 1 - 2.0 ** -53)
    lam = np.concatenate([lam, la, la[:1000], la[:1000]])
    u = np.concatenate([u, ua, np.full(1000, 2.0 ** -53), np.full(1000, 1 - 2.0 ** -53)])
    exact = np.empty(lam.size, np.int32)
    lib().oracle_poisson_inv_many(u.ctypes.data, lam.ctypes.data, lam.size, exact.ctypes.data)
    for rel in (0.0, 1.6e-6, -1.6e-6):
        k, safe = _presearch(u, lam, rel)
        shortcut = u <= (1.0 - lam) - 4.5e-16          # settled before the pre-search on host and device alike
        assert np.all(exact[shortcut] == 0)
        chk = safe & ~shortcut
        assert np.array_equal(k[chk], exact[chk]), np.flatnonzero(chk & (k != exact))[:5]
        assert np.mean(~safe[: 2 * n] & ~shortcut[: 2 * n]) < 3e-4
