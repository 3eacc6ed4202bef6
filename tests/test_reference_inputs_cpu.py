"""The synthetic population written in the reference's own input formats (beep_b200.synth.write_reference_inputs)
must stay loadable by the UNMODIFIED reference: oracle/_ref/ref_harness (built by oracle/ref/build_ref.sh where the
reference sources are present) parses the TOML/CSV files, exports the flattened model and runs its own
State::simulate on it.  Skipped where the binary is absent."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


@pytest.mark.skipif(not os.access(HARNESS, os.X_OK), reason="oracle/_ref/ref_harness not built (no reference sources here)")
def test_reference_reads_the_synthetic_inputs(tmp_path):
    from beep_b200.bpk import read_bpk
    from beep_b200.synth import make_model, write_reference_inputs
    work = str(tmp_path)
    toml = write_reference_inputs(work, A=6, G=20, T=70, name="syn6")
    out = os.path.join(work, "o.bpk")
    env = dict(os.environ, BEEP_HARNESS_BENCH="0.2")
    r = subprocess.run([HARNESS, out, "1", "3", f"inputfile={os.path.basename(toml)}", "mode=sim", "outputdir=Out"], cwd=work, env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    line = open(out + ".bench").read().split()
    kv = dict(x.split("=") for x in line[1:])
    assert (int(kv["T"]), int(kv["A"]), int(kv["D"])) == (70, 6, 20) and int(kv["reps"]) >= 1
    # the model the reference built from the files has the generator's shape and populations
    d = read_bpk(out)
    m, _ = make_model(A=6, G=20, T=70)
    assert int(d["desc.C"][0]) == m.C and int(d["desc.J"][0]) == m.J and int(d["desc.G"][0]) == 20
    pop_ref = d["desc.pop_init"].reshape(6, m.C, 20).sum(axis=1)
    pop_syn = np.asarray(m.arrays["pop_init"]).reshape(6, m.C, 20).sum(axis=1)
    assert np.array_equal(pop_ref, pop_syn)


@pytest.mark.skipif(not os.access(HARNESS, os.X_OK), reason="oracle/_ref/ref_harness not built (no reference sources here)")
def test_reference_runs_the_shipped_england_inputs(tmp_path):
    """bench.py's cpu_baseline of --workload england_as: the unmodified reference loads tests/golden/ref_inputs/COVID 19 England,
    draws a parameter vector from its prior and times its own State::simulate on the England_AS shape."""
    work = str(tmp_path)
    out = os.path.join(work, "o.bpk")
    env = dict(os.environ, BEEP_HARNESS_BENCH="0.2")
    r = subprocess.run([HARNESS, out, "1", "100", "inputfile=COVID 19 England/England_AS.toml", "mode=abcsmc", "nsample=10", f"outputdir={os.path.join(work, 'Out')}"],
                       cwd=os.path.join(ROOT, "tests", "golden", "ref_inputs"), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    kv = dict(x.split("=") for x in open(out + ".bench").read().split()[1:])
    assert (int(kv["T"]), int(kv["A"]), int(kv["D"])) == (1050, 1, 18) and int(kv["reps"]) >= 1
