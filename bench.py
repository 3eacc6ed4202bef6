#!/usr/bin/env python
"""bench.py -- throughput of the batched simulate + observation-likelihood step.

Metric (BASELINE.json): particle x area x age timesteps per second, i.e.
P*A*D*T units per step divided by the step time, whole job over all ranks.

A step is State::simulate(theta) for a batch of P particles (reference
src/state.cc:582-592): set_param (K0) + tau-leap over all T steps (K1/K2) +
observation log-likelihood (K4), on the synthetic scale-up population of
BASELINE.json configs[4] (A areas x 20 ages x 12 compartments x 365 days).
  value  inputs (theta) resident in HBM, device-timed with CUDA events;
  e2e    through the C-ABI with HOST buffers: theta copied in, L copied out.
One process per GPU (torchrun for N > 1); particles are sharded across ranks
with no data-path collective (weak scaling).

`--impl reference` times the CPU restatement of the reference algorithm
(oracle/liboracle.so, kind "port": the reference itself builds its model from
TOML/CSV files and cannot ingest the synthetic description) on all host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UNIT = 96.0     # simulate + likelihood, trajectory discarded: pop read + write, 2*C*4 B, C = 12
METRIC = "particle*area*age timesteps/sec"
UNIT = "particle*area*age*steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--areas", type=int, default=int(os.environ.get("BEEP_BENCH_AREAS", 10000)))
    ap.add_argument("--particles", type=int, default=int(os.environ.get("BEEP_BENCH_PARTICLES", 16)), help="particles per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--keep-traj", action="store_true", help="store transnum (MBP chain initialisation), 148 B/unit")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# CPU arm (oracle port) -- also the cpu_baseline leg of the GPU arm
# ---------------------------------------------------------------------------
def cpu_run(steps, warmup, budget_s=12.0):
    """Times set_param + simulate + obs log-likelihood of the oracle on every host
    core, on a bounded sample: the same synthetic generator at A = 256 areas."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from beep_b200.synth import make_model, jitter_params
    import oracle_lib
    A_s = 256
    m, theta = make_model(A=A_s, G=20, T=365)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nthread = max(1, min(cores, 64))
    th = jitter_params(theta, nthread, rel=0.05, seed=3)
    states = [oracle_lib.OracleState(m, stochastic=1, keyed=True, imap_from_pop=True) for _ in range(nthread)]
    units_one = m.T * m.A * m.D

    def one(i, rep):
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
    reps = max(1, int(budget_s / max(t1, 1e-3) / max(steps + warmup, 1)))
    reps = min(reps, 8)
    for w in range(warmup):
        for r in range(reps):
            sweep(r)
    t0 = time.perf_counter()
    for s in range(steps):
        for r in range(reps):
            sweep(100 + s * reps + r)
    dt = time.perf_counter() - t0
    units = units_one * nthread * reps * steps
    sample = (f"oracle/liboracle.so (C restatement of state.cc/obsmodel.cc), synthetic generator at A={A_s} areas x 20 ages x 365 steps, "
              f"{nthread} threads x {reps} particles per step, set_param+simulate+obs_loglik")
    return units / dt, nthread, sample, dt / steps * 1e3


def main_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    v, cores, sample, ms = cpu_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 0),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, nparam):
    return {"workload": f"synthetic scale-up (BASELINE configs[4]): {args.areas} areas x 20 ages x 12 compartments x 13 transitions x 365 steps, "
                        f"{args.particles} particles per GPU, step = set_param + tau-leap simulate + obs log-likelihood"
                        + (" (trajectory kept)" if args.keep_traj else " (trajectory discarded)"),
            "areas": args.areas, "ages": 20, "compartments": 12, "transitions": 13, "timesteps": 365,
            "particles_per_gpu": args.particles, "nparam": nparam,
            "l2": "working set (population state, > 300 MB) exceeds L2; no explicit flush"}


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
    from beep_b200.synth import make_model, jitter_params

    m, theta = make_model(A=args.areas, G=20, T=365)
    eng = E.Engine(m, local)
    P = args.particles
    flags = E.KEEP_TRANSNUM if args.keep_traj else 0
    b = eng.batch(P, flags)
    th = jitter_params(theta, P, rel=0.05, seed=100 + rank)          # each rank owns its own particles
    ids = (np.arange(P, dtype=np.uint64) + np.uint64(rank * P))
    th_dev = torch.from_numpy(th).cuda()
    th_pin = torch.from_numpy(th).pin_memory()
    Lhost = np.zeros(P)
    stream = torch.cuda.ExternalStream(b.stream, device=torch.device("cuda", local))
    units = float(P) * m.A * m.D * m.T

    def step_dev(i):
        b.set_params_dev(th_dev.data_ptr())
        b.simulate(seed=7 + i, stream_ids=None)
        E._check(b.L.beep_obs_loglik_dev(b.h, -1, None, None))

    def step_e2e(i):
        E._check(b.L.beep_set_params(b.h, th_pin.data_ptr()))
        b.simulate(seed=7 + i)
        E._check(b.L.beep_obs_loglik(b.h, -1, Lhost.ctypes.data, None))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    b.set_params_dev(th_dev.data_ptr())
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
    # the dominant kernel's own duration (CUDA events around the k_traj launch on the batch stream)
    traj_ms = []
    for i in range(max(args.steps, 1)):
        step_dev(1000 + i)
        b.sync()
        traj_ms.append(b.last_simulate_ms)
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    clk = clocks.stop() if clocks else None

    value = units * world * args.steps / (ms_dev * 1e-3)
    e2e = units * world * args.steps / (ms_e2e * 1e-3)
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
    bpu = 148.0 if args.keep_traj else BYTES_PER_UNIT
    kms = float(np.mean(traj_ms))
    achieved = units * bpu / (kms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_traj_dram_bytes_per_launch")
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, m.nparam),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(th.nbytes), "d2h_bytes_per_step": int(Lhost.nbytes),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_traj<SIM,GRID>", "kernel_ms": kms, "bytes_per_unit": bpu, "units_per_launch": units, "peak_source": peak_src},
        "clocks": clk,
    }
    if not args.no_cpu:
        v, cores, sample, _ = cpu_run(2, 1, budget_s=10.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
