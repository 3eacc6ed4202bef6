#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_harness (built by oracle/ref/build_ref.sh from the sources
under /root/reference) on the reference's own worked examples and converts the
BPK1 dumps to compressed .npz fixtures.  The GPU box has no /root/reference, so
the fixtures are committed; this script is how they were made:

    ./oracle/ref/build_ref.sh && python tests/golden/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from beep_b200.bpk import read_bpk  # noqa: E402

REF = os.environ.get("BEEP_REFERENCE", "/root/reference")
HARNESS = os.environ.get("BEEP_REF_HARNESS") or os.path.join(ROOT, "oracle", "_ref", "ref_harness")

# name -> (toml relative to the reference root, harness args, ntheta, seed)
CASES = {
    # BASELINE.json configs[0]: EX.toml in simulation mode
    "EX_sim": ("examples/EX.toml", ["mode=sim"], 1, 7),
    # the same model in inference mode (real observation values)
    "EX_inf": ("examples/EX.toml", ["mode=mcmcmbp", "nsample=10"], 3, 7),
    # BASELINE.json configs[1]
    "EX_A1": ("examples/EX_A1.toml", ["mode=mcmcmbp", "nsample=10"], 3, 11),
    # EX_B1 (marginal + population-fraction data) cannot be used: the reference itself aborts on it at this
    # commit, in inference mode ("Cannot have zero variance", data.cc:1113) and in mode=sim (Utils error 2)
    "EX_A5": ("examples/EX_A5.toml", ["mode=mcmcmbp", "nsample=10"], 2, 13),   # a second demographic classification (sex) with susceptibility
    "EX_C3": ("examples/EX_C3.toml", ["mode=mcmcmbp", "nsample=10"], 2, 27),   # areas x a demographic classification
    "EX_G3": ("examples/EX_G3.toml", ["mode=sim"], 1, 47),   # the England model structure in small (smoothed-spline priors: simulation mode, as EX_E2)
    "EX_B4": ("examples/EX_B4.toml", ["mode=mcmcmbp", "nsample=10"], 2, 17),   # Poisson observation model
    "EX_B6": ("examples/EX_B6.toml", ["mode=mcmcmbp", "nsample=10"], 2, 19),   # observation factor spline
    "EX_C2": ("examples/EX_C2.toml", ["mode=mcmcmbp", "nsample=10"], 3, 23),   # 10 areas, covariates, area effects, Dirichlet
    "EX_C4": ("examples/EX_C4.toml", ["mode=mcmcmbp", "nsample=10"], 3, 29),   # 4 age groups
    "EX_C5": ("examples/EX_C5.toml", ["mode=mcmcmbp", "nsample=10"], 2, 31),   # geographic mixing spline
    "EX_C6": ("examples/EX_C6.toml", ["mode=mcmcmbp", "nsample=10"], 2, 37),   # two strains
    "EX_D3": ("examples/EX_D3.toml", ["mode=mcmcmbp", "nsample=10"], 2, 41),   # level effects / tv covariates
    # EX_E2 has smoothed-spline priors that only mode=map can sample (model.cc:741): simulation mode
    "EX_E2": ("examples/EX_E2.toml", ["mode=sim"], 1, 43),   # time-varying age matrix
    # BASELINE.json configs[2]/[3]: the real model shape (1 area x 18 ages x 9 compartments x 10 transitions x
    # 1050 steps, 178 parameters, time-varying age matrix, ~11.5k observations).  The file as shipped does not load
    # at this commit (obsmodel="scale" is rejected, inputs.cc:467): every table is switched to "normal 10%" in a
    # scratch copy -- any England_AS parity claim carries this substitution.
    "England_AS": ("COVID 19 England/England_AS.toml", ["mode=abcsmc", "nsample=10"], 2, 53),
}
# Monte-Carlo replicate fixtures: the reference's own std:: samplers, 10 000 simulations each
STAT = {
    "stat_EX": ("examples/EX.toml", ["mode=sim"], 1, 61, 10000),
    "stat_EX_C4": ("examples/EX_C4.toml", ["mode=sim"], 1, 67, 10000),
}
POSTERIOR = {"post_EX_A1": ("examples/EX_A1.toml", 2000)}      # BASELINE.json configs[1]: ~2.5 minutes of the reference
ABCSMC = {"abcsmc_EX_A1": ("examples/EX_A1.toml", 2000, 6),     # ~3 minutes of the reference
          # England_AS (BASELINE configs[2]; obsmodel scale -> normal 10%, see SUBST): generation 0 only = 2000 draws from the prior,
          # simulated, with their error functions.  Later generations are out of reach: at this size the reference's own
          # run does not leave generation 1 (Cholesky of the 170-parameter proposal kernel) within half an hour of CPU.
          "abcsmc_England_AS": ("COVID 19 England/England_AS.toml", 2000, 1)}
# the reference's own mode=pmcmc (PMCMC::run with its particle filter on the CPU): posterior estimates for the adapter's run
# of the SAME driver on the engine.  invT is fixed (left unset the reference tunes it during burn-in, and two runs then
# target different power posteriors).  ~2 minutes of the reference.
PMCMC_POST = {"pmcmc_EX_A1": ("examples/EX_A1.toml", ["nparticle=32", "nsample=2000", "invT=0.03"])}
# the reference's own mode=map (ML::run: CMA-ES over State::likelihood_approx, ml.cc:150-385): best log posterior and step size
# per generation, final estimates and model evidence, for the adapter's run of the SAME driver with the generation's
# likelihoods coming from beep_likelihood_approx
MAP_POST = {"map_EX_A1": ("examples/EX_A1.toml", ["nparticle=48", "ngeneration=20", "nsample=50", "seed=7"])}
DECISIONS = {"pmcmc": ["mode=pmcmc", "nparticle=24", "nsample=1000"], "abcmbp": ["mode=abcmbp", "nparticle=24", "ngeneration=3"],
             "pas": ["mode=pas", "nparticle=12", "nrun=2", "invT_final=1"], "mc3": ["mode=mc3", "nchain=6", "nrun=2", "invT_final=1", "nsample=1000"]}
MVN_CASES = {
    "mvn_small": (12, 7, 60, 16, 71),          # a handful of free parameters (EX_A1-like)
    "mvn_england": (180, 140, 400, 24, 73),    # England_AS-like: 180 parameters, most of them free
}
# observation-model branches no example exercises (plain normal, negative binomial, censored and unknown data):
# the loaded observations of these cases are re-labelled by the harness (BEEP_HARNESS_OBSPATCH, ref_harness.cc)
OBSPATCH = {"OBS_C4": ("examples/EX_C4.toml", ["mode=mcmcmbp", "nsample=10"], 2, 83),
            "OBS_B4": ("examples/EX_B4.toml", ["mode=mcmcmbp", "nsample=10"], 2, 89)}
# the MVN-approximate likelihood of mode=map (State::likelihood_approx, state_mvn_approx.cc:11-91) with ML's own set-up
# (BEEP_HARNESS_APPROX = invT, ref_harness.cc): L for three parameter vectors (the TOML values, two prior draws), the mean
# populations and carried variances of the first
APPROX = {"APX_A1": ("examples/EX_A1.toml", ["mode=map"], 3, 101),    # one area, one group
          "APX_B4": ("examples/EX_B4.toml", ["mode=map"], 3, 103),    # Poisson observation model, marginal data
          "APX_C4": ("examples/EX_C4.toml", ["mode=map"], 3, 107),    # four age groups
          "APX_C2": ("examples/EX_C2.toml", ["mode=map"], 3, 109),    # ten areas, covariates
          "APX_D3": ("examples/EX_D3.toml", ["mode=map"], 3, 113),    # level effects
          "APX_C5": ("examples/EX_C5.toml", ["mode=map"], 3, 127)}    # 32 areas coupled by the geographic mixing matrix (128 compartments)
# ObservationModel::get_graph_state (obsmodel.cc:113-170), the series behind the Posterior/state files: the harness adds the
# graphs' points of the deterministic and the stochastic trajectory (BEEP_HARNESS_GRAPHS); descriptions carry the graph fields
GRAPHS = {"GRS_C4": ("examples/EX_C4.toml", ["mode=mcmcmbp", "nsample=10"], 2, 131),    # four age groups: a graph per group
          "GRS_C2": ("examples/EX_C2.toml", ["mode=mcmcmbp", "nsample=10"], 2, 137),    # ten areas
          "GRS_B6": ("examples/EX_B6.toml", ["mode=mcmcmbp", "nsample=10"], 2, 139),    # observation factor spline
          "GRS_B4": ("examples/EX_B4.toml", ["mode=mcmcmbp", "nsample=10"], 2, 149)}    # marginal data
# model modifications of mode=prediction (model.modelmod, model.cc:1440-1585; state.cc:415,446-453): the reference constructs its
# Model in prediction mode from TOML files extended IN THE SCRATCH TREE with 'modification' entries of every kind, and the harness
# records its own set_param / simulate / Mbp::mbp under them (no posterior samples are needed for that)
PRED = {"PRED_F3": ("pred_F3.toml", ["mode=prediction"], 2, 151),      # the reference's own scenario example: a transition rate doubled from day 80
        "PRED_C2": ("pred_C2.toml", ["mode=prediction"], 2, 157)}      # ten areas: beta, external force, spline factor, spline set, per-area and per-transition filters
# Mbp::get_al (mbp.cc:334-361) in both modes on a real proposal (arrays mbp.al.* of the harness); a fixture of its own so that
# the older ones stay byte for byte what they were
ACCEPT = {"AL_A1": ("examples/EX_A1.toml", ["mode=mcmcmbp", "nsample=10"], 3, 163), "AL_C4": ("examples/EX_C4.toml", ["mode=mcmcmbp", "nsample=10"], 3, 167)}
SUBST = {"England_AS": [('obsmodel="scale"', 'obsmodel="normal 10%"')]}
BIG = {"England_AS"}


def convert(bpk_path, out_path):
    d = read_bpk(bpk_path)
    out = {}
    for k, v in d.items():
        if ".approx." in k and not k.startswith("th0."):
            continue                        # the state arrays of the first parameter vector are enough
        if k.endswith("det.transmean"):
            # deterministic dynamics: transnum IS transmean (state.cc:627)
            assert np.array_equal(v, d[k.replace("transmean", "transnum")])
            continue
        if k.endswith(".transnum") and ".det." not in k:
            assert np.all(v == np.round(v))
            v = v.astype(np.int32)
        if k.endswith(".mask"):
            v = v.astype(np.uint8)
        out[k] = v
    np.savez_compressed(out_path, **out)
    return len(out)


def write_dcc_inputs(work):
    ex = os.path.join(REF, "examples")
    frac = "time,Male,Female\n0,0.5,0.5\n40,0.62,0.38\n90,0.45,0.55\n140,0.3,0.7\n"
    for name in ("C3", "C2"):
        ddir = os.path.join(work, "Data_dcc_" + name)
        shutil.copytree(os.path.join(ex, "Data_EX_" + name), ddir)
        open(os.path.join(ddir, "sexfrac.csv"), "w").write(frac)
        txt = open(os.path.join(ex, f"EX_{name}.toml")).read()
        txt = txt.replace(f'datadir = "examples/Data_EX_{name}"', f'datadir = "Data_dcc_{name}"')
        txt = txt.replace(f'outputdir = "examples/Output_EX_{name}"', f'outputdir = "Output_dcc_{name}"')
        if name == "C2":
            # a second classification for the ten areas: Male / Female counts next to the population column
            rows = open(os.path.join(ddir, "population.csv")).read().strip().split("\n")
            out = [rows[0] + ",Male,Female"]
            for k, r in enumerate(rows[1:]):
                n = int(r.split(",")[1])
                male = int(n * (0.44 + 0.01 * k))
                out.append(f"{r},{male},{n - male}")
            open(os.path.join(ddir, "population.csv"), "w").write("\n".join(out) + "\n")
            txt = txt.replace('comps = [', 'democats = [{name="Sex", cats="Male | Female"}]\n\ncomps = [', 1)
        txt += '\ndemocat_change = [{name="Sex", file="sexfrac.csv"}]\n'
        open(os.path.join(work, f"dcc_{name}.toml"), "w").write(txt)


def write_pred_inputs(work):
    ex = os.path.join(REF, "examples")
    txt = open(os.path.join(ex, "EX_F3.toml")).read()
    txt = txt.replace('modification = [{type="trans_rate_fac", start="80", trans="S->E", factor="2.0"}]',
                      'modification = [{type="trans_rate_fac", start="80", trans="S->E", factor="2.0"},\n{type="trans_rate_fac", start="40", end="60", trans="I->R", factor="0.7"},\n{type="efoi_fac", start="20", end="50", factor="3.0"},\n{type="spline_fac", name="R_t", start="85", factor="0.5"},\n{type="spline_set", name="R_t", start="60", end="75", value="1.3"}]')
    open(os.path.join(work, "pred_F3.toml"), "w").write(txt)
    txt = open(os.path.join(ex, "EX_C2.toml")).read()
    assert "prediction_end" not in txt and "\nmodification" not in txt
    end = [ln for ln in txt.splitlines() if ln.startswith("end")][0]
    t_end = int(end.split('"')[1])
    txt = txt.replace(end, end + f'\nprediction_end = "{t_end + 30}"\nmodification = [{{type="beta_fac", start="{t_end // 2}", factor="0.6"}},\n'
                      f'{{type="efoi_fac", start="10", end="{t_end // 3}", factor="2.5"}},\n{{type="trans_rate_fac", start="{t_end - 20}", trans="I->R", factor="1.4"}}]', 1)
    open(os.path.join(work, "pred_C2.toml"), "w").write(txt)


def main():
    only = sys.argv[1:]
    if not os.path.exists(HARNESS):
        subprocess.check_call([os.path.join(ROOT, "oracle", "ref", "build_ref.sh")])
    work = tempfile.mkdtemp(prefix="beep_golden_")
    # the reference resolves datadir/outputdir relative to the cwd and writes
    # its output there: run in a scratch tree of symlinks, never in /root/reference
    os.makedirs(os.path.join(work, "examples"))
    for f in os.listdir(os.path.join(REF, "examples")):
        os.symlink(os.path.join(REF, "examples", f), os.path.join(work, "examples", f))
    eng = os.path.join(work, "COVID 19 England")
    os.makedirs(eng)
    os.symlink(os.path.join(REF, "COVID 19 England", "Data_England_AS"), os.path.join(eng, "Data_England_AS"))
    for name, subs in SUBST.items():
        toml = CASES[name][0]
        txt = open(os.path.join(REF, toml)).read()
        for old, new in subs:
            txt = txt.replace(old, new)
        open(os.path.join(work, toml), "w").write(txt)
    # the synthetic scale-up shape (BASELINE.json configs[4]) at a size the reference dumps in seconds, written
    # in the reference's own input formats by beep_b200/synth.py: 12 areas x 20 ages x 12 compartments x 140 steps
    from beep_b200.synth import write_reference_inputs
    write_reference_inputs(work, A=12, G=20, T=140, name="synthetic12")
    CASES["SYN_A12"] = ("synthetic12.toml", ["mode=sim"], 1, 59)
    BIG.add("SYN_A12")
    CASES.update(OBSPATCH)
    CASES.update(APPROX)
    CASES.update(GRAPHS)
    CASES.update(ACCEPT)
    write_pred_inputs(work)
    CASES.update(PRED)
    # demographic change (data.democat_change, state.cc:339-376): no input of the reference uses it, so two of its
    # examples are extended IN THE SCRATCH TREE -- EX_C3 (one area, Sex) and EX_C2 (ten areas, Sex added to its
    # population file; the change applies to all of them) -- with a table of category fractions over time
    write_dcc_inputs(work)
    CASES["DCC_C3"] = ("dcc_C3.toml", ["mode=mcmcmbp", "nsample=10"], 2, 61)
    CASES["DCC_C2"] = ("dcc_C2.toml", ["mode=mcmcmbp", "nsample=10"], 2, 67)
    # DCC_C2 comes from the reference's sources compiled -O3 -ffp-contract=off (oracle/_ref/nofma, same recipe, sources
    # untouched).  With several areas a demographic change couples every area's population to every other's, so in a mixed
    # SIMU / MBP proposal the means of the MBP areas differ from the initial ones by rounding-level amounts, and whether a
    # cell consumes a draw from the SEQUENTIAL stream (mbp.cc:286-311: rate_p >, ==, < rate_i) then depends on the last
    # bit: the reference's own flags (-mfma, contraction on) fuse pop += frac_dif * sum (state.cc:368) where the oracle
    # rounds twice.  Every other quantity of the fixture agrees with the contracted build to 1e-9 (checked once by hand).
    nofma = os.path.join(ROOT, "oracle", "_ref", "nofma")
    if (not only or "DCC_C2" in only) and not os.path.exists(os.path.join(nofma, "ref_harness")):
        subprocess.check_call([os.path.join(ROOT, "oracle", "ref", "build_ref.sh")], env=dict(os.environ, BEEP_REF_OUT=nofma, BEEP_REF_OPT="-O3 -ffp-contract=off"))
    for name, (toml, args, ntheta, seed) in CASES.items():
        if only and name not in only:
            continue
        bpk = os.path.join(work, name + ".bpk")
        cmd = [os.path.join(nofma, "ref_harness") if name == "DCC_C2" else HARNESS, bpk, str(ntheta), str(seed), f"inputfile={toml}"] + args
        env = dict(os.environ)
        if name in OBSPATCH:
            env["BEEP_HARNESS_OBSPATCH"] = "1"
        if name in BIG:
            env["BEEP_HARNESS_BIG"] = "1"      # skip pop / Imap / Idiag dumps
        if name in APPROX:
            env["BEEP_HARNESS_APPROX"] = "1"   # invT = 1
        if name in GRAPHS:
            env["BEEP_HARNESS_GRAPHS"] = "1"
        r = subprocess.run(cmd, cwd=work, capture_output=True, text=True, env=env)
        if r.returncode != 0 or not os.path.exists(bpk):
            print(f"{name}: harness FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        out = os.path.join(ROOT, "tests", "golden", name + ".npz")
        n = convert(bpk, out)
        if name in ACCEPT:                  # only the acceptance arrays are kept
            dd = dict(np.load(out))
            dd = {k: v for k, v in dd.items() if k.startswith("mbp.al.")}
            np.savez_compressed(out, **dd)
            n = len(dd)
        print(f"{name}: {n} arrays, {os.path.getsize(out) / 1024:.0f} KiB")
    for name, (toml, args, ntheta, seed, reps) in STAT.items():
        if only and name not in only:
            continue
        bpk = os.path.join(work, name + ".bpk")
        cmd = [os.path.join(nofma, "ref_harness") if name == "DCC_C2" else HARNESS, bpk, str(ntheta), str(seed), f"inputfile={toml}"] + args
        env = dict(os.environ, BEEP_HARNESS_REPS=str(reps), BEEP_HARNESS_BIG="1")
        r = subprocess.run(cmd, cwd=work, capture_output=True, text=True, env=env)
        if r.returncode != 0 or not os.path.exists(bpk):
            print(f"{name}: harness FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        d = read_bpk(bpk)
        out = {k: v for k, v in d.items() if k.startswith("desc.") or k in ("theta", "meta.seed")}
        for k in ("stat.trans_total", "stat.pop_final"):
            assert np.all(d[k] == np.round(d[k]))
            out[k] = d[k].astype(np.int32)
        out["stat.EF"] = d["stat.EF"].astype(np.float32)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {reps} replicates, {os.path.getsize(path) / 1024:.0f} KiB")
    # posterior fixtures: the reference's own MCMC-MBP run (mode=mcmcmbp, invT=1) -> Parameter_estimates.csv
    for name, (toml, nsample) in POSTERIOR.items():
        if only and name not in only:
            continue
        outdir = os.path.join(work, "out_" + name)
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "BEEP"), f"inputfile={toml}", "mode=mcmcmbp", f"nsample={nsample}", "invT=1", f"outputdir={outdir}"],
                           cwd=work, capture_output=True, text=True)
        est = os.path.join(outdir, "Parameter_estimates.csv")
        if not os.path.exists(est):
            print(f"{name}: reference run FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        import csv
        import json
        rows = list(csv.DictReader(open(est)))
        out = {"command": f"BEEP inputfile={toml} mode=mcmcmbp nsample={nsample} invT=1", "parameters": [
            {"name": x["Name"], "mean": float(x["Mean"]), "sd": float(x["SD"]), "ess": (float(x["ESS"]) if x["ESS"] not in ("---", "") else None), "prior": x["Prior"]} for x in rows]}
        json.dump(out, open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w"), indent=1)
        print(f"{name}: {len(rows)} parameters")
        # one of the run's posterior sample files (State::save, state.cc:709-739) as the fixture of the sample-file round trip
        sdir = os.path.join(ROOT, "tests", "golden", "samples_" + name[len("post_"):])
        os.makedirs(sdir, exist_ok=True)
        for suffix in ("_parameter.csv", "_transition.csv"):
            shutil.copy(os.path.join(outdir, "Posterior", "samples", "sample7" + suffix), os.path.join(sdir, "sample7" + suffix))
    # ABC-SMC end to end (BASELINE's second metric; abcsmc.cc:26-131): the reference's own run on EX_A1 -- cut-off and
    # posterior means per generation (Diagnostics/Generation.csv), final estimates, the final samples' error functions
    for name, (toml, nsample, ngen) in ABCSMC.items():
        if only and name not in only:
            continue
        import csv
        import json
        outdir = os.path.join(work, "out_" + name)
        args = [f"inputfile={toml}", "mode=abcsmc", f"nsample={nsample}", f"ngeneration={ngen}", "cutoff_frac=0.5", f"outputdir={outdir}"]
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "BEEP")] + args, cwd=work, capture_output=True, text=True)
        gfile = os.path.join(outdir, "Diagnostics", "Generation.csv")
        if not os.path.exists(gfile):
            print(f"{name}: reference run FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        lines = [ln for ln in open(gfile) if not ln.startswith("#")]
        gens = list(csv.DictReader(lines))
        est = list(csv.DictReader(open(os.path.join(outdir, "Parameter_estimates.csv"))))
        ef = []
        sdir = os.path.join(outdir, "Posterior", "samples")
        for f in sorted(os.listdir(sdir)):
            if f.endswith("_parameter.csv"):
                ef.append(float(open(os.path.join(sdir, f)).read().split("Error function,")[1].split()[0]))
        gen0 = [ln for ln in r.stdout.splitlines() if ln.startswith("Generation 0")]
        out = {"command": "BEEP " + " ".join(args[:-1]),
               "EFcut0": float(gen0[0].split("EF Cut-off:")[1].split()[0]) if gen0 else None,
               "generations": [{k: (float(v) if v not in ("UNSET", "") else None) for k, v in g.items() if k.endswith("(mean)") or k in ("Generation", "EF cut-off", "Model evidence")} for g in gens],
               "estimates": [{"name": x["Name"], "mean": float(x["Mean"]), "sd": float(x["SD"]), "prior": x["Prior"]} for x in est],
               "final_EF": sorted(ef)}
        json.dump(out, open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w"), indent=1)
        print(f"{name}: {len(gens)} generations, {len(ef)} samples")
    for name, (toml, args) in MAP_POST.items():
        if only and name not in only:
            continue
        import csv
        import json
        outdir = os.path.join(work, "out_" + name)
        cmd = [f"inputfile={toml}", "mode=map"] + args
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "BEEP")] + cmd + [f"outputdir={outdir}"], cwd=work, capture_output=True, text=True)
        est = os.path.join(outdir, "Parameter_estimates.csv")
        if not os.path.exists(est):
            print(f"{name}: reference run FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        gens = []
        for ln in r.stdout.splitlines():
            if ln.startswith("Generation ") and "Best log(Post. prob.):" in ln:
                gens.append({"best": float(ln.split("Best log(Post. prob.):")[1].split()[0]), "sigma": float(ln.split("Sigma:")[1].split()[0])})
        ev = [ln for ln in r.stdout.splitlines() if ln.startswith("The log of the model evidence is")]
        out = {"command": "BEEP " + " ".join(cmd), "generations": gens, "evidence": float(ev[0].split()[-1]) if ev else None,
               "parameters": [{"name": x["Name"], "mean": float(x["Mean"]), "sd": float(x["SD"]), "prior": x["Prior"]} for x in csv.DictReader(open(est))]}
        json.dump(out, open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w"), indent=1)
        print(f"{name}: {len(gens)} generations")
    for name, (toml, args) in PMCMC_POST.items():
        if only and name not in only:
            continue
        import csv
        import json
        outdir = os.path.join(work, "out_" + name)
        cmd = [f"inputfile={toml}", "mode=pmcmc"] + args
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "BEEP")] + cmd + [f"outputdir={outdir}"], cwd=work, capture_output=True, text=True)
        est = os.path.join(outdir, "Parameter_estimates.csv")
        if not os.path.exists(est):
            print(f"{name}: reference run FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        rows = list(csv.DictReader(open(est)))
        out = {"command": "BEEP " + " ".join(cmd), "parameters": [
            {"name": x["Name"], "mean": float(x["Mean"]), "sd": float(x["SD"]), "ess": (float(x["ESS"]) if x["ESS"] not in ("---", "") else None), "prior": x["Prior"]} for x in rows]}
        json.dump(out, open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w"), indent=1)
        print(f"{name}: {len(rows)} parameters")
    # resampling decisions by the reference's own PMCMC::bootstrap, ABCMBP::EF_cutoff,
    # PAS::bootstrap and MC3::swap_states with ran() injected at link time (oracle/ref/ref_decisions.cc)
    if not only or "decisions" in only:
        dec = {}
        for which, args in DECISIONS.items():
            bpk = os.path.join(work, f"dec_{which}.bpk")
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_decisions"), which, bpk, "5", "inputfile=examples/EX_A1.toml", f"outputdir={work}/dec_out"] + args,
                               cwd=work, capture_output=True, text=True)
            if r.returncode != 0 or not os.path.exists(bpk):
                print(f"decisions/{which}: FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
                continue
            for k, v in read_bpk(bpk).items():
                dec[f"{which}.{k}"] = v
        path = os.path.join(ROOT, "tests", "golden", "decisions.npz")
        np.savez_compressed(path, **dec)
        print(f"decisions: {len(dec)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
    # ABC-SMC weight kernel: the reference's MVN::setup / propose / probability
    # (oracle/ref/ref_mvn.cc): name -> (nparam, nvar, Nlast, Nprop, seed)
    for name, (nparam, nvar, nl, npr, seed) in MVN_CASES.items():
        if only and name not in only:
            continue
        bpk = os.path.join(work, name + ".bpk")
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_mvn"), bpk, str(nparam), str(nvar), str(nl), str(npr), str(seed)],
                           cwd=work, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(bpk):
            print(f"{name}: ref_mvn FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **read_bpk(bpk))
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
