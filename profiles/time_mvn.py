#!/usr/bin/env python
"""Times beep_mvn_kernel_sums (K7, ABC-SMC weights) at a generation's size through the C ABI with host
buffers (copies inside the timed region).  usage: python profiles/time_mvn.py [N] [Nlast] [nvar]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from beep_b200 import engine as E
from beep_b200.model import ModelDesc
from test_mvn_gpu import _random_case
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
Nl = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
nvar = int(sys.argv[3]) if len(sys.argv) > 3 else 170
m = ModelDesc.from_prefixed(dict(np.load(os.path.join(ROOT, "tests", "golden", "EX_A1.npz"))))
eng = E.Engine(m)
case = _random_case(np.random.default_rng(1), nvar + 10, nvar, Nl, N)
E.mvn_kernel_sums(eng, *case)
t0 = time.perf_counter(); reps = 5
for _ in range(reps): s = E.mvn_kernel_sums(eng, *case)
dt = (time.perf_counter() - t0) / reps
print(f"N={N} Nlast={Nl} nvar={nvar}: {dt*1e3:.2f} ms per call (host buffers, Cholesky + copies included), "
      f"{N*Nl/dt/1e9:.2f} G pairs/s, {3*N*Nl*nvar/dt/1e12:.2f} TFLOP/s FP64 (sub+fma per pair and variable); "
      f"reference double loop would be {N*Nl*nvar*nvar*3/1e12:.1f} TFLOP")
