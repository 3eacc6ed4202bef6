"""Posterior parity with the reference (BASELINE.json north_star, third criterion; configs[1] EX_A1.toml).

The reference's own MCMC-MBP run (`BEEP inputfile=examples/EX_A1.toml mode=mcmcmbp nsample=2000 invT=1`,
tests/golden/make_golden.py -> tests/golden/post_EX_A1.json: posterior mean, sd and effective sample size of
the two free parameters) against the C++ host driver host/mbp_mcmc.cc, which runs 512 independent chains (6 000 burn-in sweeps from over-dispersed starts, 2 000 sampled sweeps) of
model-based proposals through the C ABI (beep_set_params, beep_mbp, beep_obs_loglik, beep_accept).
Tolerance: the reference's Monte-Carlo error -- 4 standard errors of its mean (sd / sqrt(ESS)) -- for the
means, 25 % for the standard deviations (the relative error of an sd estimate at ESS ~ 150 is ~ 6 %)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def test_mbp_posterior_matches_reference(tmp_path):
    from beep_b200.bpk import write_bpk
    from beep_b200.model import ModelDesc
    exe = os.path.join(ROOT, "host", "mbp_mcmc")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    m, d = load_golden("EX_A1")
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    packed = ModelDesc(arrays).to_prefixed()
    packed["theta"] = np.ascontiguousarray(d["theta"][0], dtype=np.float64)
    path = str(tmp_path / "EX_A1.bpk")
    write_bpk(path, packed)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "post_EX_A1.json")))["parameters"]
    # parameter 1 = mean occupancy of I, Uniform(1,10); parameter 3 = R, Uniform(0.5,4) (EX_A1.toml)
    out = subprocess.run([exe, path, "512", "6000", "2000", "7", "1:1:10", "3:0.5:4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    got = {}
    for ln in out.stdout.splitlines():
        f = ln.split()
        if f and f[0] == "param":
            got[int(f[1])] = (float(f[3]), float(f[5]), float(f[7]))
    print(out.stdout)
    for k, r in ((1, ref[0]), (3, ref[2])):
        mean, sd, acc = got[k]
        se = r["sd"] / np.sqrt(r["ess"])
        assert abs(mean - r["mean"]) < 4 * se, (k, mean, r)
        assert abs(sd - r["sd"]) < 0.25 * r["sd"], (k, sd, r)
        assert 0.1 < acc < 0.7
