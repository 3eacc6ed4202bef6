#!/usr/bin/env python
"""One batch of beep_likelihood_approx for profiling under ncu.  usage: python profiles/prof_approx.py CASE P"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beep_b200 import engine as E
from beep_b200.model import ModelDesc
name, P = sys.argv[1], int(sys.argv[2])
d = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
m = ModelDesc.from_prefixed(d)
eng = E.Engine(m)
rng = np.random.default_rng(3)
th = np.repeat(d["theta"][:1], P, axis=0)
free = np.abs(d["theta"][1] - d["theta"][0]) > 0
th[:, free] *= 1 + 0.03 * rng.standard_normal((P, int(free.sum())))
bt = eng.batch(P)
bt.set_params(th)
print(bt.likelihood_approx(1.0)[:4])
