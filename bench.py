#!/usr/bin/env python
"""bench.py -- throughput of BEEP's batched latent-state hot path on B200.

Metric (BASELINE.json): particle x area x age timesteps per second, i.e. P*A*D*T units per step
divided by the step time, whole job over all ranks.

Default workload (N = 1 and N > 1): the synthetic scale-up population of BASELINE.json configs[4]
(10 000 areas x 20 ages x 12 compartments x 13 transitions x 365 steps; 16 particles per GPU, the
(A, P) point on the memory frontier of SURVEY.md 8d).  A step is State::simulate(theta) for the
batch (reference src/state.cc:582-592): set_param (K0) + tau-leap over all T steps (K1/K2) +
observation log-likelihood (K4).
  value  inputs (theta) resident in HBM, device-timed with CUDA events on the engine's stream;
  e2e    through the C-ABI with HOST buffers: theta copied in, L copied out, every step.
Other workloads (extra lines for DESIGN.md, not the driver's headline):
  --workload england_as   the real model shape of configs[2]/[3] (1 area x 18 ages x 1050 steps,
                          ~11.5k observations) from tests/golden/England_AS.npz, 16 384 particles;
  --mode mbp              one model-based proposal sweep (Mbp::mbp, mbp.cc:252-330) + likelihood.
One process per GPU (torchrun for N > 1); particles are sharded across ranks with no data-path
collective (weak scaling).

Every line also carries two extra objects measured AFTER the headline (they are not part of `value`):
  "filter"   BASELINE configs[3]: the bootstrap particle filter of PMCMC::obs_prob (pmcmc.cc:180-268) on the
             England_AS model, 4 096 particles in all sharded over the ranks (strong scaling) -- sections
             simulated on the device, weights all-gathered (ncclAllGather), the distinct remote ancestors moved
             HBM to HBM (beep_end_sec_swap), one trajectory drawn (beep_particle_sample); seconds per filter
             pass and the collectives' share of it;
  "abcsmc"   BASELINE's second metric: seconds per ABC-SMC generation (abcsmc.cc:26-131) of 16 384 England_AS
             simulations in all, run by the reference's OWN host code on the engine (host/_ref/BEEP_b200:
             TOML inputs, priors, MVN proposals of the reference; batched engine calls; one process per GPU,
             samples exchanged by beep_allgather_scalar), with the reference's mode=abcsmc simulations/s beside it.

  "map"      row N3: State::likelihood_approx of mode=map (state_mvn_approx.cc:11-91), the objective ML's CMA-ES and
             Hessian code evaluate for many parameter vectors (ml.cc:389-413, 540-575), batched: beep_likelihood_approx on
             the EX_C5 model (32 areas, 128 compartments, 400 steps, 1 792 observations; fixture APX_C5), 592 parameter
             vectors per GPU through the C ABI with host buffers; the CPU restatement on one host core beside it.

`--impl reference` times the UNMODIFIED reference on all host cores (kind "reference"): oracle/_ref/ref_harness
(the reference's own objects, built here from /root/reference and shipped as a binary) runs the synthetic
population written in the reference's own TOML/CSV formats (--workload england_as: the England_AS inputs shipped under
tests/golden/ref_inputs, one draw from the prior per process).  Where that is not possible (binary absent, MBP mode) it
falls back to the CPU restatement of the reference algorithm (oracle/liboracle.so, kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle*area*age timesteps/sec"
UNIT = "particle*area*age*steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BEEP_BENCH_WORKLOAD", "synthetic"), choices=["synthetic", "england_as"])
    ap.add_argument("--mode", default="simulate", choices=["simulate", "mbp"])
    ap.add_argument("--areas", type=int, default=int(os.environ.get("BEEP_BENCH_AREAS", 10000)))
    ap.add_argument("--particles", type=int, default=int(os.environ.get("BEEP_BENCH_PARTICLES", 0)), help="particles per GPU (0 = workload default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--keep-traj", action="store_true", help="store transnum (MBP chain initialisation)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)     # one launch under ncu (see measure_traffic)
    ap.add_argument("--no-extras", action="store_true", help="skip the filter / abcsmc objects")
    a = ap.parse_args()
    if a.particles <= 0:
        a.particles = 16 if a.workload == "synthetic" else 16384
        if a.mode == "mbp" and a.workload == "synthetic":
            a.particles = 4
    return a


# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
def load_workload(args, areas=None):
    """(ModelDesc, theta_true, description)"""
    from beep_b200.synth import make_model
    if args.workload == "synthetic":
        A = areas if areas is not None else args.areas
        nn = int(os.environ.get("BEEP_BENCH_NNEIGH", 8))      # neighbours per area (8: the sparse matrix of SURVEY 8d; A-1: a dense one)
        m, theta = make_model(A=A, G=20, T=365, nneigh=min(nn, A - 1))
        return m, theta
    from beep_b200.model import ModelDesc
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "England_AS.npz")))
    m = ModelDesc.from_prefixed(d)
    return m, d["theta"][0]


def particle_params(args, theta, P, seed):
    from beep_b200.synth import jitter_params
    if args.workload == "synthetic":
        return jitter_params(theta, P, rel=0.05, seed=seed)
    return np.repeat(theta[None, :], P, axis=0)          # prior-sampled vector of the fixture; bounds are model specific


def bytes_per_unit(m, mode, keep):
    """Algorithmic HBM bytes per (particle, area, dp, step) unit, SURVEY.md 8(d): counts int32, means FP64."""
    C, J = m.C, m.J
    if mode == "mbp":
        return J * 4 + J * 8 + J * 4 + J * 8 + 2 * C * 4      # read n_i, lambda_i; write n_p, lambda_p; pop r/w
    return 2 * C * 4 + (J * 4 if keep else 0)


def workload_config(args, m, nparam):
    if args.workload == "synthetic":
        what = f"synthetic scale-up (BASELINE configs[4]): {m.A} areas x {m.D} ages x {m.C} compartments x {m.J} transitions x {m.T} steps"
        l2 = "state per step (population tiles, > 300 MB at 10 000 areas x 16 particles) exceeds L2; no explicit flush"
    else:
        what = (f"England_AS model shape (BASELINE configs[2]/[3], obsmodel scale->normal 10%): {m.A} area x {m.D} ages x {m.C} compartments x "
                f"{m.J} transitions x {m.T} steps, {m.nobs} observations")
        l2 = "population tiles stay in shared memory for the whole trajectory; per-particle tables (> 1 GB over the batch) exceed L2"
    step = "set_param + tau-leap simulate + obs log-likelihood" if args.mode == "simulate" else "set_param + model-based proposal sweep (Mbp::mbp) + obs log-likelihood"
    return {"workload": f"{what}, {args.particles} particles per GPU, step = {step}" + (" (trajectory kept)" if args.keep_traj or args.mode == "mbp" else " (trajectory discarded)"),
            "areas": int(m.A), "ages": int(m.D), "compartments": int(m.C), "transitions": int(m.J), "timesteps": int(m.T),
            "particles_per_gpu": args.particles, "nparam": int(nparam), "mode": args.mode, "l2": l2}


# ---------------------------------------------------------------------------
# CPU arm (oracle port) -- also the cpu_baseline leg of the GPU arm
# ---------------------------------------------------------------------------
def cpu_run(args, steps, warmup, budget_s=12.0):
    """Times the oracle's set_param + simulate (or mbp) + obs log-likelihood on every host core, on a
    bounded sample of the workload (synthetic: the same generator at 256 areas; England_AS: as is)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    m, theta = load_workload(args, areas=256)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nthread = max(1, min(cores, 64))
    th = particle_params(args, theta, nthread, seed=3)
    states = [oracle_lib.OracleState(m, stochastic=1, keyed=True, imap_from_pop=True) for _ in range(nthread)]
    props = None
    if args.mode == "mbp":
        props = [oracle_lib.OracleState(m, stochastic=1, keyed=True, imap_from_pop=True) for _ in range(nthread)]
        for i in range(nthread):
            states[i].set_param(th[i])
            states[i].simulate(seed=1, stream_id=i)
    units_one = m.T * m.A * m.D

    def one(i, rep):
        if args.mode == "mbp":
            props[i].set_param(th[i] * (1 + 0.01 * (np.arange(th.shape[1]) % 7 == 3)))
            props[i].mbp(states[i], None, seed=1000 + rep, stream_id=i)
            props[i].obs_loglik()
        else:
            st = states[i]
            st.set_param(th[i])
            st.simulate(seed=1000 + rep, stream_id=i)
            st.obs_loglik()

    def sweep(rep):
        ts = [threading.Thread(target=one, args=(i, rep)) for i in range(nthread)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    t1 = sweep(0)
    reps = max(1, min(8, int(budget_s / max(t1, 1e-3) / max(steps + warmup, 1))))
    for w in range(warmup):
        for r in range(reps):
            sweep(r)
    t0 = time.perf_counter()
    for s in range(steps):
        for r in range(reps):
            sweep(100 + s * reps + r)
    dt = time.perf_counter() - t0
    units = units_one * nthread * reps * steps
    what = "set_param+mbp+obs_loglik" if args.mode == "mbp" else "set_param+simulate+obs_loglik"
    shape = f"synthetic generator at A={m.A} areas x {m.D} ages x {m.T} steps" if args.workload == "synthetic" else "England_AS shape"
    sample = (f"oracle/liboracle.so (C restatement of state.cc/mbp.cc/obsmodel.cc), {shape}, {nthread} threads x {reps} particles per step, {what}")
    return units / dt, nthread, sample, dt / steps * 1e3, int(m.A)


REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def reference_areas(nproc):
    """The reference keeps pop / transnum / transmean for all T steps per State (about 2.3 MB per area at the
    synthetic shape) and reads the geographic matrix only as a dense CSV (data_matrix.cc:206-229: A^2 numbers),
    so the 10 000-area configuration itself cannot be loaded; the reference arm runs the same generator at the
    largest area count that fits a third of the free host memory over `nproc` processes, at most 2 048."""
    free = 64e9
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                free = float(ln.split()[1]) * 1024
    except Exception:
        pass
    a = int(free / 3 / nproc / 5.5e6)          # measured: about 4-5 MB of resident memory per area and process
    return int(max(256, min(2048, a // 64 * 64)))


def reference_run(args, steps, budget_s=4.0, areas=None):
    """Times the UNMODIFIED reference (oracle/_ref/ref_harness: the reference's own objects, built from the
    sources under /root/reference by oracle/ref/build_ref.sh and shipped as a binary) on the synthetic
    population written in the reference's own TOML/CSV formats (beep_b200.synth.write_reference_inputs):
    one single-rank process per host core -- what the reference's MPI ranks do between collectives -- each
    looping State::simulate(param) + ObservationModel::calculate with the reference's own samplers.
    Returns (units/s, cores, sample, ms_per_step, areas) or None when the binary is absent or the workload has no
    reference input files."""
    england = args.workload == "england_as"
    ref_inputs = os.path.join(ROOT, "tests", "golden", "ref_inputs")
    if args.mode != "simulate" or not os.access(REF_HARNESS, os.X_OK) or (england and not os.path.isdir(os.path.join(ref_inputs, "COVID 19 England"))):
        return None
    import tempfile
    import shutil
    from beep_b200.synth import write_reference_inputs
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nproc = max(1, min(cores, 64))
    if areas is None and not england:
        areas = reference_areas(nproc)
    work = tempfile.mkdtemp(prefix="beep_ref_")
    try:
        env = dict(os.environ, BEEP_HARNESS_BENCH=str(budget_s * max(steps, 1)))
        if england:
            # England_AS as shipped under tests/golden/ref_inputs (obsmodel scale -> normal 10%, see its README): every process
            # simulates its own draw from the prior (the reference's Model::sample_from_prior, seed 100 + i), as ABC-SMC does
            areas = 1
            procs = [subprocess.Popen(["timeout", "-s", "KILL", "900", REF_HARNESS, os.path.join(work, f"o{i}.bpk"), "1", str(100 + i), "inputfile=COVID 19 England/England_AS.toml",
                                       "mode=abcsmc", "nsample=10", f"outputdir={os.path.join(work, f'Output_{i}')}"],
                                      cwd=ref_inputs, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(nproc)]
        else:
            toml = write_reference_inputs(work, A=areas)
            procs = [subprocess.Popen(["timeout", "-s", "KILL", "900", REF_HARNESS, os.path.join(work, f"o{i}.bpk"), "1", str(100 + i), f"inputfile={os.path.basename(toml)}", "mode=sim", f"outputdir=Output_{i}"],
                                      cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(nproc)]
        units, secs, ok = 0.0, 0.0, 0
        for i, p in enumerate(procs):
            rc = p.wait(timeout=1000)
            res = os.path.join(work, f"o{i}.bpk.bench")
            if rc != 0 or not os.path.exists(res):
                print(f"bench.py: reference process {i} failed (rc={rc})", file=sys.stderr)
                continue
            kv = dict(x.split("=") for x in open(res).read().split()[1:])
            units += float(kv["reps"]) * float(kv["T"]) * float(kv["A"]) * float(kv["D"])
            secs = max(secs, float(kv["seconds"]))
            ok += 1
        if ok == 0:
            return None
        nproc = ok
    except Exception as e:
        print(f"bench.py: reference arm falls back to the CPU port: {e!r}", file=sys.stderr)
        return None
    finally:
        shutil.rmtree(work, ignore_errors=True)
    if england:
        sample = (f"oracle/_ref/ref_harness (the unmodified reference's State::simulate + ObservationModel::calculate, its own std:: samplers) on England_AS "
                  f"(tests/golden/ref_inputs, obsmodel scale -> normal 10%), one draw from the prior per process, {nproc} single-rank processes x {secs:.1f} s")
        return units / secs, nproc, sample, secs / max(steps, 1) * 1e3, areas
    sample = (f"oracle/_ref/ref_harness (the unmodified reference's State::simulate + ObservationModel::calculate, its own std:: samplers), "
              f"synthetic population in the reference's TOML/CSV formats at A={areas} areas x 20 ages x 365 steps (the largest the reference's dense "
              f"geographic-matrix loader and per-State trajectory storage allow on this host), {nproc} single-rank processes x {secs:.1f} s")
    return units / secs, nproc, sample, secs / max(steps, 1) * 1e3, areas


def main_reference(args):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    kind = "reference"
    r = reference_run(args, args.steps + args.warmup)
    if r is None:
        kind = "port"
        r = cpu_run(args, args.steps, args.warmup)
    v, cores, sample, ms, areas = r
    m, theta = load_workload(args, areas=areas)
    cfg = workload_config(args, m, theta.size)
    if args.workload == "synthetic":
        cfg["areas"] = int(m.A)                     # what the reference arm actually ran (see reference_areas)
        cfg["workload"] += f" [reference arm: {int(m.A)} areas, not {args.areas}: the reference cannot load the dense geographic matrix of the full size]"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



# ---------------------------------------------------------------------------
# DRAM traffic of the dominant kernel, measured in this run
# ---------------------------------------------------------------------------
def traffic_child(args):
    """One launch of the bench's trajectory kernel (run under ncu by measure_traffic)."""
    from beep_b200 import engine as E
    m, theta = load_workload(args)
    eng = E.Engine(m, 0)
    mbp = args.mode == "mbp"
    flags = (E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN) if mbp else (E.KEEP_TRANSNUM if args.keep_traj else 0)
    b = eng.batch(args.particles, flags)
    b.set_params(particle_params(args, theta, args.particles, seed=100))
    b.simulate(seed=7)
    b.sync()


def measure_traffic(args, local):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the trajectory kernel at the bench configuration:
    a child process of this script under `ncu --metrics ... -k regex:k_traj -c 1` on this rank's GPU.  None when
    ncu is not there or fails (the traffic is reported as measured or not at all)."""
    import shutil
    ncu = shutil.which("ncu")
    if ncu is None or args.mode != "simulate":
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:k_traj", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", args.workload, "--mode", args.mode, "--areas", str(args.areas),
           "--particles", str(args.particles)] + (["--keep-traj"] if args.keep_traj else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = str(local) if "CUDA_VISIBLE_DEVICES" not in os.environ else os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env).stdout
        import csv
        import io
        tot, seen = 0.0, 0
        lines = out.splitlines()
        start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
        for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
            if row.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(row["Metric Value"].replace(",", ""))
                unit = row.get("Metric Unit", "byte").lower()
                tot += v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
                seen += 1
        return tot if seen == 2 else None
    except Exception as e:
        print(f"bench.py: traffic measurement failed: {e!r}", file=sys.stderr)
        return None


# ---------------------------------------------------------------------------
# extra objects: the sharded particle filter (configs[3]) and ABC-SMC seconds per generation (the second metric)
# ---------------------------------------------------------------------------
ADAPTER = os.path.join(ROOT, "host", "_ref", "BEEP_b200")
REF_BEEP = os.path.join(ROOT, "oracle", "_ref", "BEEP")
REF_INPUTS = os.path.join(ROOT, "tests", "golden", "ref_inputs")


def filter_extra(rank, world, local, dist, ntot=4096, reps=3):
    """PMCMC::obs_prob (pmcmc.cc:180-268) on England_AS, ntot particles over `world` GPUs (strong scaling)."""
    import torch
    from beep_b200 import engine as E
    from beep_b200.model import ModelDesc
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "England_AS.npz")))
    arrays = dict(ModelDesc.from_prefixed(d).arrays)
    arrays["stochastic"] = 1
    m = ModelDesc(arrays)
    eng = E.Engine(m, local)
    uid = None
    if world > 1:
        box = [E.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    cm = E.Comm(eng, uid, world, rank)
    N = ntot // world
    b = eng.batch(N, E.KEEP_TRANSNUM)
    theta = np.repeat(d["theta"][:1], N, axis=0)
    ids = np.arange(rank * N, (rank + 1) * N, dtype=np.uint64)
    nsec = eng.nsection

    def one(seed):
        tcomm = 0.0
        b.set_params(theta)
        back = np.zeros((nsec, ntot), dtype=np.uint32)
        ll = 0.0
        for s in range(nsec):
            ti, tf = eng.section_range(s)
            if s > 0:
                t0 = time.perf_counter()
                if world > 1:
                    cm.end_sec_swap(b, back[s - 1])
                else:
                    b.permute(back[s - 1], E.PERMUTE_SECTION_STATE)
                tcomm += time.perf_counter() - t0
            b.simulate(ti, tf, seed=seed, stream_ids=ids if s == 0 else None)
            L = b.obs_loglik(section=s)
            t0 = time.perf_counter()
            Lall = cm.allgather(L)
            tcomm += time.perf_counter() - t0
            inc, back[s] = E.bootstrap(Lall, 1.0, seed, s)
            ll += inc
        t0 = time.perf_counter()
        cm.particle_sample(b, back)
        tcomm += time.perf_counter() - t0
        return ll, tcomm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one(1)
    barrier()
    t0 = time.perf_counter()
    tc = 0.0
    for r in range(reps):
        ll, c = one(2 + r)
        tc += c
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, tc], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, tc = float(t[0]), float(t[1])
    sent, recv = cm.bytes_moved
    out = {"workload": f"bootstrap particle filter (PMCMC::obs_prob) on England_AS, {ntot} particles in all over {world} GPU(s), {nsec} sections, resampling after each, one trajectory sampled",
           "seconds_per_filter": dt / reps, "particle_steps_per_s": ntot * m.T * m.D * reps / dt, "collectives_share": tc / dt,
           "collectives": "ncclAllGather of section log-likelihoods + beep_end_sec_swap (distinct remote ancestors, HBM to HBM) + beep_particle_sample" if world > 1 else "one GPU: device permutation only",
           "bytes_sent_per_rank_per_filter": sent / (reps + 1), "log_likelihood": ll, "scaling": "strong"}
    cm.close()
    b.close()
    eng.close()
    return out


def map_extra(local, P=592, reps=3, with_cpu=True):
    """likelihoods/s of the MVN-approximate likelihood (mode=map) on one GPU; see the module docstring"""
    from beep_b200 import engine as E
    from beep_b200.model import ModelDesc
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "APX_C5.npz")))
    m = ModelDesc.from_prefixed(d)
    eng = E.Engine(m, local)
    rng = np.random.default_rng(3)
    th = np.repeat(d["theta"][:1], P, axis=0)
    free = np.abs(d["theta"][1] - d["theta"][0]) > 0
    th[:, free] *= 1 + 0.03 * rng.standard_normal((P, int(free.sum())))
    bt = eng.batch(P)
    bt.set_params(th)
    L = bt.likelihood_approx(1.0)
    t0 = time.perf_counter()
    for _ in range(reps):
        bt.set_params(th)
        L = bt.likelihood_approx(1.0)
    dt = (time.perf_counter() - t0) / reps
    out = {"workload": "EX_C5 (32 areas x 4 compartments = 128 MVN components, 400 steps, 1792 observations), State::likelihood_approx at invT=1",
           "vectors": P, "seconds_per_batch": dt, "likelihoods_per_s": P / dt, "finite": int(np.isfinite(L).sum())}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import OracleState
        o = OracleState(m)
        t0 = time.perf_counter()
        Lo = o.likelihood_approx(th[0], 1.0)
        tc = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / tc, "unit": "likelihoods/s", "cores": 1, "kind": "port", "sample": "one evaluation of the CPU restatement (oracle/beep_oracle_approx.c)",
                               "rel_diff_vs_engine": abs(L[0] - Lo) / abs(Lo)}
    bt.close()
    return out


def abcsmc_extra(rank, world, local, nonce, ntot=16384):
    """Seconds per ABC-SMC generation (abcsmc.cc:26-131) of ntot England_AS simulations in all, by the reference's own
    host code on the engine (host/_ref/BEEP_b200), one process per GPU.  Generation 0 (prior samples): with the
    reference's proposal kernel over all 170 free parameters next to none of the later generations' proposals fall
    inside the prior bounds -- in the reference itself as well -- so they measure the host's rejection loop."""
    if not os.access(ADAPTER, os.X_OK):
        return None
    env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local), BEEP_BATCH=str(ntot // world),
               BEEP_NCCL_ID_FILE=f"/tmp/beep_nccl_id_{os.environ.get('MASTER_PORT', '0')}_{nonce}")
    cmd = ["timeout", "-s", "KILL", "300", ADAPTER, "inputfile=COVID 19 England/England_AS.toml", "mode=abcsmc", f"nsample={ntot // 2}", "ngeneration=1",
           "cutoff_frac=0.5", f"outputdir=/tmp/beep_b200_abcsmc_{rank}"]
    r = subprocess.run(cmd, cwd=REF_INPUTS, capture_output=True, text=True, env=env)
    if rank != 0:
        return None
    gen = [ln for ln in r.stdout.splitlines() if ln.startswith("generation 0:")]
    if r.returncode != 0 or not gen:
        print(f"bench.py: ABC-SMC adapter failed: {r.stdout[-500:]} {r.stderr[-500:]}", file=sys.stderr)
        return None
    kv = dict(x.split("=") for x in gen[0].split(":", 1)[1].split())
    return {"workload": f"ABC-SMC generation 0 on England_AS (obsmodel scale -> normal 10%): {int(float(kv['simulations']))} simulations in all over {world} GPU(s): prior samples, "
                        "State::simulate, error function, cut-off at cutoff_frac = 0.5, samples exchanged over the ranks; host code = the reference's own objects (host/_ref/BEEP_b200)",
            "s_per_generation": float(kv["seconds"]), "simulations_per_s": float(kv["simulations/s"]), "EFcut": float(kv["EFcut"])}


def abcsmc_reference(nsample=24):
    """The reference's own mode=abcsmc generation 0 on the host cores: one single-rank process per core, each 2 nsample
    simulations; start-up (input parsing) measured by a run of 2 simulations and subtracted."""
    if not os.access(REF_BEEP, os.X_OK):
        return None
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nproc = max(1, min(cores, 64))

    def run(ns, n):
        t0 = time.perf_counter()
        ps = [subprocess.Popen(["timeout", "-s", "KILL", "600", REF_BEEP, "inputfile=COVID 19 England/England_AS.toml", "mode=abcsmc", f"nsample={ns}", "ngeneration=1", "cutoff_frac=0.5",
                                f"outputdir=/tmp/beep_ref_abcsmc_{i}", f"seed={i}"], cwd=REF_INPUTS, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(n)]
        ok = all(p.wait() == 0 for p in ps)
        return time.perf_counter() - t0, ok

    t_start, ok0 = run(1, nproc)
    t_full, ok1 = run(nsample, nproc)
    if not (ok0 and ok1) or t_full <= t_start:
        return None
    sims = nproc * 2 * (nsample - 1)
    return {"simulations_per_s": sims / (t_full - t_start), "cores": nproc, "kind": "reference",
            "sample": f"oracle/_ref/BEEP mode=abcsmc generation 0 on England_AS, {nproc} single-rank processes x {2 * nsample} simulations, start-up ({t_start:.1f} s) subtracted"}

# ---------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def main_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from beep_b200 import engine as E

    m, theta = load_workload(args)
    eng = E.Engine(m, local)
    P = args.particles
    mbp = args.mode == "mbp"
    keep = args.keep_traj or mbp
    flags = (E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN) if mbp else (E.KEEP_TRANSNUM if args.keep_traj else 0)
    b = eng.batch(P, flags)
    th = particle_params(args, theta, P, seed=100 + rank)          # each rank owns its own particles
    ids = (np.arange(P, dtype=np.uint64) + np.uint64(rank * P))
    Lhost = np.zeros(P)
    units = float(P) * m.A * m.D * m.T
    init = None
    if mbp:
        # the chains' current states; the timed step proposes from them (Mbp::mbp) and evaluates the likelihood
        init = b
        init.set_params(th)
        init.simulate(0, m.T, seed=1, stream_ids=ids)
        b = eng.batch(P, flags)
        th = th * (1 + 0.01 * (np.arange(th.shape[1]) % 7 == 3))
    th_dev = torch.from_numpy(th).cuda()
    th_pin = torch.from_numpy(th).pin_memory()
    stream = torch.cuda.ExternalStream(b.stream, device=torch.device("cuda", local))

    def step_dev(i):
        b.set_params_dev(th_dev.data_ptr())
        if mbp:
            b.mbp(init, None, seed=7 + i)
        else:
            b.simulate(seed=7 + i)
        E._check(b.L.beep_obs_loglik_dev(b.h, -1, None, None))

    def step_e2e(i):
        E._check(b.L.beep_set_params(b.h, th_pin.data_ptr()))
        if mbp:
            b.mbp(init, None, seed=7 + i)
        else:
            b.simulate(seed=7 + i)
        E._check(b.L.beep_obs_loglik(b.h, -1, Lhost.ctypes.data, None))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    b.set_params_dev(th_dev.data_ptr())
    if mbp:
        b.mbp(init, None, seed=1, stream_ids=ids)
    else:
        b.simulate(0, m.T, seed=1, stream_ids=ids)       # installs the stream ids; first warm-up
    b.sync()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count
        ev0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        ev1.record(stream)
        barrier()
        launches = eng.launch_count - l0
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    clocks = Clocks(local) if rank == 0 else None
    ms_dev, launches = timed(step_dev, args.steps, args.warmup)
    # the dominant kernel's own duration (CUDA events around the k_traj launch on the batch's stream)
    traj_ms = []
    for i in range(max(args.steps, 1)):
        step_dev(1000 + i)
        b.sync()
        traj_ms.append(b.last_simulate_ms)
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    clk = clocks.stop() if clocks else None

    value = units * world * args.steps / (ms_dev * 1e-3)
    e2e = units * world * args.steps / (ms_e2e * 1e-3)
    extra_filter = extra_abcsmc = extra_map = None
    if not args.no_extras and args.workload == "synthetic" and args.mode == "simulate":
        if rank == 0:
            try:
                extra_map = map_extra(local, with_cpu=not args.no_cpu)
            except Exception as e:
                print(f"bench.py: map extra failed: {e!r}", file=sys.stderr)
        try:
            extra_filter = filter_extra(rank, world, local, dist)
        except Exception as e:
            print(f"bench.py: filter extra failed on rank {rank}: {e!r}", file=sys.stderr)
        try:
            box = [os.urandom(6).hex()]
            if dist is not None:
                dist.broadcast_object_list(box, src=0)          # a fresh name for the ncclUniqueId hand-over file
            extra_abcsmc = abcsmc_extra(rank, world, local, box[0])
        except Exception as e:
            print(f"bench.py: abcsmc extra failed on rank {rank}: {e!r}", file=sys.stderr)
        if dist is not None:
            dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    bpu = float(bytes_per_unit(m, args.mode, keep))
    kms = float(np.mean(traj_ms))
    achieved = units * bpu / (kms * 1e-3) / 1e9
    traffic = measure_traffic(args, local)
    sync = eng.sync_scope
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, m, m.nparam),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(th.nbytes), "d2h_bytes_per_step": int(Lhost.nbytes),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": (f"k_traj_chunk<{'MBP' if mbp else 'SIM'},NT={eng.chunk_tiles}>" if eng.chunk_tiles else f"k_traj<{'MBP' if mbp else 'SIM'},{sync}>"),
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, measured in this run" if traffic is not None else None, "kernel_ms": kms, "bytes_per_unit": bpu, "units_per_launch": units,
                     "peak_source": peak_src},
        "clocks": clk,
    }
    if not args.no_cpu:
        kind = "reference"
        r = reference_run(args, 3)
        if r is None:
            kind = "port"
            r = cpu_run(args, 2, 1, budget_s=10.0)
        v, cores, sample, _, _ = r
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if extra_filter is not None:
        line["filter"] = extra_filter
    if extra_abcsmc is not None:
        if not args.no_cpu and world == 1:
            extra_abcsmc["cpu_baseline"] = abcsmc_reference()
        line["abcsmc"] = extra_abcsmc
    if extra_map is not None:
        line["map"] = extra_map
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.traffic_child:
        traffic_child(a)
    elif a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
