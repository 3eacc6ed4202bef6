"""The reference's own host code on the engine (host/_ref/BEEP_b200, built by host/ref_adapter/build.sh from
the unmodified reference objects + host/ref_adapter/beep_b200_main.cc) against the reference's own runs.

ABC-SMC (BASELINE's second metric, abcsmc.cc:26-131) on EX_A1: same TOML, same nsample / ngeneration /
cutoff_frac; the fixture tests/golden/abcsmc_EX_A1.json holds the reference's cut-off and posterior means per
generation and the error functions of its final samples (tests/golden/make_golden.py).  Stochastic parity:
cut-offs within 12 %, posterior means within a quarter of the posterior sd, and the two-sample Kolmogorov-Smirnov
distance between the final samples' error functions (relative to the cut-off) below 0.1 -- the distance, not the
p-value: the realised cut-off sequence and proposal kernels differ from run to run of the SAME program, so two
runs' samples are not draws from one distribution (two seeds of the adapter are compared the same way).
PMCMC (BASELINE configs[3]'s driver) on EX_A1: the reference's own PMCMC::run with its likelihood estimates coming from
the engine's particle filter, posterior against the reference's own mode=pmcmc run.  England_AS: generation 0 of
ABC-SMC against the reference's.  Needs the prebuilt binary (it travels with the repo)."""
import csv
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "host", "_ref", "BEEP_b200")
INPUTS = os.path.join(GOLDEN, "ref_inputs")


def _run(args, env=None, timeout=600):
    if not os.path.exists(EXE):
        pytest.skip("host/_ref/BEEP_b200 has not been built (needs the reference sources: host/ref_adapter/build.sh)")
    r = subprocess.run(["timeout", "-s", "KILL", str(timeout), EXE] + args, cwd=INPUTS, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_abcsmc_matches_reference_run(tmp_path):
    from scipy import stats
    ref = json.load(open(os.path.join(GOLDEN, "abcsmc_EX_A1.json")))
    out = str(tmp_path / "out")
    _run(["inputfile=examples/EX_A1.toml", "mode=abcsmc", "nsample=2000", "ngeneration=6", "cutoff_frac=0.5", f"outputdir={out}", "seed=3"])
    gens = list(csv.DictReader(open(os.path.join(out, "b200_generations.csv"))))
    assert len(gens) == 6
    cut = [float(g["EFcut"]) for g in gens]
    ref_cut = [ref["EFcut0"]] + [g["EF cut-off"] for g in ref["generations"]]
    for a, b in zip(cut, ref_cut):
        assert abs(a - b) <= 0.12 * b, (cut, ref_cut)
    rows = list(csv.DictReader(open(os.path.join(out, "b200_posterior_parameters.csv"))))
    w = np.array([float(r["weight"]) for r in rows])
    ef = np.array([float(r["EF"]) for r in rows])
    assert np.all(ef < cut[-1])
    for est in ref["estimates"]:
        x = np.array([float(r[est["name"]]) for r in rows])
        mean = float(np.sum(w * x) / np.sum(w))
        if est["sd"] == 0:
            assert mean == est["mean"]
        else:
            assert abs(mean - est["mean"]) < 0.25 * est["sd"], (est, mean)
    ks = stats.ks_2samp(ef / cut[-1], np.array(ref["final_EF"]) / ref_cut[-1])
    assert ks.statistic < 0.1, ks
    # the same comparison between two seeds of the adapter itself: the yardstick for the line above
    out2 = str(tmp_path / "out2")
    _run(["inputfile=examples/EX_A1.toml", "mode=abcsmc", "nsample=2000", "ngeneration=6", "cutoff_frac=0.5", f"outputdir={out2}", "seed=11"])
    rows2 = list(csv.DictReader(open(os.path.join(out2, "b200_posterior_parameters.csv"))))
    cut2 = float(list(csv.DictReader(open(os.path.join(out2, "b200_generations.csv"))))[-1]["EFcut"])
    ks2 = stats.ks_2samp(ef / cut[-1], np.array([float(r["EF"]) for r in rows2]) / cut2)
    print(f"KS distance adapter vs reference {ks.statistic:.3f} (p = {ks.pvalue:.2g}); adapter seed 3 vs seed 11 {ks2.statistic:.3f} (p = {ks2.pvalue:.2g})")
    assert ks2.statistic < 0.1


def test_filter_runs_through_the_adapter():
    out = _run(["inputfile=examples/EX_A1.toml", "mode=pmcmc", "nparticle=512", "nsample=10", "outputdir=/tmp/beep_b200_filter_out"],
               env={"BEEP_RUN": "filter", "BEEP_FILTER_REPS": "2"})
    lines = [ln for ln in out.splitlines() if ln.startswith("filter ")]
    assert len(lines) == 2
    ll = [float(ln.split("log-likelihood=")[1].split()[0]) for ln in lines]
    assert all(np.isfinite(ll)) and abs(ll[0] - ll[1]) < 10          # two seeds of the same filter
    assert all(float(ln.split("sampled_events=")[1]) > 0 for ln in lines)


def test_pmcmc_posterior_matches_reference_run(tmp_path):
    """mode=pmcmc: the reference's OWN PMCMC driver (PMCMC::run, pmcmc.cc:33-110 -- burn-in, ParamProp proposals and their
    adaptation, Metropolis-Hastings, trace and posterior files, object code unchanged) with PMCMC::obs_prob and
    Mpi::particle_sample replaced at link time by the engine's particle filter and ancestry walk (host/ref_adapter).
    Same TOML, nparticle, nsample and fixed invT as the reference's own run behind tests/golden/pmcmc_EX_A1.json:
    posterior means agree within 4 standard errors (sd / sqrt(ESS), both runs), posterior sds within 35 %."""
    ref = json.load(open(os.path.join(GOLDEN, "pmcmc_EX_A1.json")))
    out = str(tmp_path / "out")
    args = ref["command"].split()[1:]
    log = _run(args + [f"outputdir={out}"], timeout=900)
    assert "particle filters of 32 particles" in log
    rows = {r["Name"]: r for r in csv.DictReader(open(os.path.join(out, "Parameter_estimates.csv")))}
    msg = []
    for est in ref["parameters"]:
        r = rows[est["name"]]
        mean, sd = float(r["Mean"]), float(r["SD"])
        if est["sd"] == 0:
            assert mean == est["mean"]
            continue
        ess = float(r["ESS"])
        se = np.hypot(sd / np.sqrt(ess), est["sd"] / np.sqrt(est["ess"]))
        msg.append(f"{est['name']}: {mean:.3f} +- {sd:.3f} (ESS {ess:.0f}) vs reference {est['mean']:.3f} +- {est['sd']:.3f} (ESS {est['ess']:.0f})")
        assert abs(mean - est["mean"]) < 4 * se, msg[-1]
        assert abs(sd - est["sd"]) < 0.35 * est["sd"], msg[-1]
    print("; ".join(msg) + "; " + [ln for ln in log.splitlines() if ln.startswith("pmcmc on the engine")][0])


def test_abcsmc_england_prior_predictive(tmp_path):
    """England_AS (BASELINE configs[2]/[3]; obsmodel scale -> normal 10 %, tests/golden/ref_inputs/README.md), the real
    model shape end to end: generation 0 of ABC-SMC = draws from the prior (the reference's own Model::sample_from_prior
    in both programs), State::simulate, error function, cut-off at cutoff_frac.  The accepted error functions of the
    adapter on the engine and of the reference's own run (fixture tests/golden/abcsmc_England_AS.json, 2 000 accepted
    samples) are two independent samples of ONE distribution: two-sample Kolmogorov-Smirnov below the alpha = 0.001
    critical distance, and the cut-offs (medians of the prior-predictive error function) agree.  Later generations of
    the reference are out of reach on England_AS (see tests/golden/make_golden.py)."""
    from scipy import stats
    ref = json.load(open(os.path.join(GOLDEN, "abcsmc_England_AS.json")))
    out = str(tmp_path / "out")
    _run(["inputfile=COVID 19 England/England_AS.toml", "mode=abcsmc", "nsample=2000", "ngeneration=1", "cutoff_frac=0.5", f"outputdir={out}", "seed=5"], timeout=900)
    gens = list(csv.DictReader(open(os.path.join(out, "b200_generations.csv"))))
    cut, ref_cut = float(gens[0]["EFcut"]), ref["EFcut0"]
    assert abs(cut - ref_cut) <= 2e-5 * ref_cut, (cut, ref_cut)
    rows = list(csv.DictReader(open(os.path.join(out, "b200_posterior_parameters.csv"))))
    ef = np.array([float(r["EF"]) for r in rows])
    ref_ef = np.array(ref["final_EF"])
    assert ef.size == ref_ef.size == 2000
    lim = min(cut, ref_cut)                      # both samples truncated at the same point
    a, b = ef[ef < lim], ref_ef[ref_ef < lim]
    assert a.size > 1800 and b.size > 1800
    ks = stats.ks_2samp(a, b)
    crit = 1.95 * np.sqrt((a.size + b.size) / (a.size * b.size))
    print(f"England_AS generation 0: cut-off {cut:.2f} vs {ref_cut:.2f}; KS distance {ks.statistic:.4f} (critical {crit:.4f} at alpha = 0.001, p = {ks.pvalue:.3g}); "
          f"deciles of cut - EF: {np.quantile(lim - a, [0.1, 0.5, 0.9]).round(2)} vs {np.quantile(lim - b, [0.1, 0.5, 0.9]).round(2)}")
    assert ks.statistic < crit, ks


def test_map_matches_reference_run(tmp_path):
    """mode=map (row N3): the reference's OWN ML driver (ML::run -> cmaes, ml.cc:150-385: sampling, ranking, covariance and step-size
    adaptation, Hessian, posterior samples, outputs -- object code unchanged) with ML::generate_samples (ml.cc:389-413) replaced at
    link time by ONE beep_likelihood_approx call per generation.  Same TOML, seed, nparticle and ngeneration as the reference's own
    run behind tests/golden/map_EX_A1.json.  The sampler's ran() stream is the reference's own and the engine's likelihoods
    agree with State::likelihood_approx to 1e-12, so the two CMA-ES runs follow the same path: best log posterior and step size of
    every generation agree to the printed digits (1e-5 relative), and so do the final estimates and the model evidence."""
    ref = json.load(open(os.path.join(GOLDEN, "map_EX_A1.json")))
    out = str(tmp_path / "out")
    log = _run(ref["command"].split()[1:] + [f"outputdir={out}"], timeout=600)
    gens = []
    for ln in log.splitlines():
        if ln.startswith("Generation ") and "Best log(Post. prob.):" in ln:
            gens.append((float(ln.split("Best log(Post. prob.):")[1].split()[0]), float(ln.split("Sigma:")[1].split()[0])))
    assert len(gens) == len(ref["generations"]) == 20
    for (best, sigma), g in zip(gens, ref["generations"]):
        assert abs(best - g["best"]) <= 1e-5 * abs(g["best"]), (gens, ref["generations"])
        assert abs(sigma - g["sigma"]) <= 1e-5 * g["sigma"]
    rows = {r["Name"]: r for r in csv.DictReader(open(os.path.join(out, "Parameter_estimates.csv")))}
    for est in ref["parameters"]:
        assert abs(float(rows[est["name"]]["Mean"]) - est["mean"]) <= 2e-3 * max(abs(est["mean"]), 1e-9), est
        assert abs(float(rows[est["name"]]["SD"]) - est["sd"]) <= 2e-2 * max(est["sd"], 1e-9) + 1e-3, est
    ev = [ln for ln in log.splitlines() if ln.startswith("The log of the model evidence is")]
    assert ev and abs(float(ev[0].split()[-1]) - ref["evidence"]) <= 1e-3 * abs(ref["evidence"])
    line = [ln for ln in log.splitlines() if ln.startswith("map on the engine")][0]
    assert "20 generations of 48 approximate likelihoods" in line
    print(line)
