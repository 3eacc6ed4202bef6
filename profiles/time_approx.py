#!/usr/bin/env python
"""Times beep_likelihood_approx (row N3: State::likelihood_approx of mode=map, batched) through the C ABI with host
buffers, next to the CPU restatement on one host core.  usage: python profiles/time_approx.py [case ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from beep_b200 import engine as E
from beep_b200.model import ModelDesc
from oracle_lib import OracleState
cases = sys.argv[1:] or ["APX_A1", "APX_C4", "APX_C2", "APX_C5"]
for name in cases:
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    m = ModelDesc.from_prefixed(d)
    eng = E.Engine(m)
    o = OracleState(m)
    t0 = time.perf_counter(); Lo = o.likelihood_approx(d["theta"][0], 1.0); tc = time.perf_counter() - t0
    n = m.A * m.C * m.D
    for P in (148, 592, 2368):
        rng = np.random.default_rng(3)
        th = np.repeat(d["theta"][:1], P, axis=0)
        free = np.abs(d["theta"][1] - d["theta"][0]) > 0
        th[:, free] *= 1 + 0.03 * rng.standard_normal((P, int(free.sum())))
        bt = eng.batch(P)
        bt.set_params(th)
        L = bt.likelihood_approx(1.0)
        t0 = time.perf_counter(); reps = 3
        for _ in range(reps):
            bt.set_params(th); L = bt.likelihood_approx(1.0)
        dt = (time.perf_counter() - t0) / reps
        print(f"{name}: n={n} T={m.T} nobs={m.nobs} P={P}: {dt*1e3:.1f} ms per batch = {dt/P*1e6:.1f} us per likelihood "
              f"({P/dt:.0f} /s); CPU restatement {tc*1e3:.1f} ms per likelihood on one core ({1/tc:.0f} /s): x{tc*P/dt:.0f}", flush=True)
        bt.close()
