"""The collectives of the C-ABI library (beep_b200/csrc/comm.cuh: NCCL all-gather, Mpi::copy_particles,
Mpi::end_sec_swap, Mpi::particle_sample -- reference src/mpi.cc:81-165,215-389) on real GPUs.

One-GPU tests run everywhere; the multi-rank ones need at least two GPUs (`gpurun --gpus 2 ...`): one process
per GPU, the ncclUniqueId handed over through a file, every call through the C ABI (engine.Comm)."""
import os
import sys
import tempfile
import time

import numpy as np
import pytest

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu
UNSET = 0xFFFFFFFF


def _walk(trajs, back, sections):
    """Mpi::particle_sample (mpi.cc:331-389) on host arrays: trajs[p] is particle p's transnum"""
    nsec = len(sections)
    out = np.zeros_like(trajs[0])
    p = int(back[nsec - 1][0])
    for sec in range(nsec - 1, -1, -1):
        ti, tf = sections[sec]
        out[ti:tf] = trajs[p][ti:tf]
        if sec > 0:
            p = int(back[sec - 1][p])
    return out


@pytest.mark.parametrize("name", ["EX_A1", "EX_C4", "EX_C2"])
def test_particle_sample_one_gpu(name):
    from beep_b200 import engine as E
    from beep_b200.shard import ShardedFilter
    from beep_b200.model import ModelDesc
    m, d = load_golden(name)
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    ms = ModelDesc(arrays)
    eng = E.Engine(ms)
    N = 48
    b = eng.batch(N, E.KEEP_TRANSNUM)
    f = ShardedFilter(b)
    f.run(d["theta"][0], seed=9)
    back = np.stack(f.ancestry)
    sections = [eng.section_range(s) for s in range(eng.nsection)]
    trajs = [b.trajectory(k) for k in range(N)]
    th, tn = E.particle_sample(b, back)
    assert np.array_equal(tn, _walk(trajs, back, sections))
    assert np.array_equal(th, d["theta"][0])
    # a communicator of one rank behaves like none
    cm = E.Comm(eng, None, 1, 0)
    th2, tn2 = cm.particle_sample(b, back)
    assert np.array_equal(tn2, tn)
    assert np.array_equal(cm.allgather(np.arange(5.0)), np.arange(5.0))
    src = np.full(N, UNSET, dtype=np.uint32)
    src[3] = 7
    cm.copy_particles(b, src)
    assert np.array_equal(b.trajectory(3), trajs[7])


def _worker(rank, world, name, idfile, out):
    import torch
    from beep_b200 import engine as E
    from beep_b200.shard import ShardedFilter
    from beep_b200.model import ModelDesc
    torch.cuda.set_device(rank)
    m, d = load_golden(name)
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    ms = ModelDesc(arrays)
    eng = E.Engine(ms, rank)
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(E.comm_unique_id())
        os.replace(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.01)
    uid = open(idfile, "rb").read()
    cm = E.Comm(eng, uid, world, rank)
    N = 96 // world
    ntot = N * world
    res = {}
    # all-gather
    g = cm.allgather(np.arange(N, dtype=np.float64) + 1000 * rank)
    res["gather"] = g.tolist()
    # particle filter: section likelihoods all-gathered, end_sec_swap between sections
    b = eng.batch(N, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    f = ShardedFilter(b, rank, world, comm=cm)
    ll = f.run(d["theta"][0], seed=11)
    back = np.stack(f.ancestry)
    res["ll"] = ll
    res["back"] = back.tolist()
    ps = cm.particle_sample(b, back)
    if rank == 0:
        res["sample"] = ps[1].tolist()
    res["swap_bytes"] = cm.bytes_moved
    # copy_particles: a decision with local copies, remote copies, chains and keeps
    rng = np.random.default_rng(5)
    theta = np.repeat(d["theta"][:1], N, axis=0) * (1 + 0.01 * rng.uniform(-1, 1, (ntot, d["theta"].shape[1])))[rank * N:(rank + 1) * N]
    b.set_params(theta)
    b.simulate(seed=21, stream_ids=np.arange(rank * N, (rank + 1) * N, dtype=np.uint64))
    L0 = cm.allgather(b.obs_loglik())
    Li0 = cm.allgather(b.latent_loglik())
    sums0 = cm.allgather(np.array([b.trajectory(k).sum() for k in range(N)]))
    pc = np.full(ntot, UNSET, dtype=np.uint32)
    idx = rng.choice(ntot, size=ntot // 2, replace=False)
    pc[idx] = rng.integers(0, ntot, size=idx.size)
    cm.copy_particles(b, pc)
    L1 = cm.allgather(b.obs_loglik())
    Li1 = cm.allgather(b.latent_loglik())
    sums1 = cm.allgather(np.array([b.trajectory(k).sum() for k in range(N)]))
    src = np.where(pc == UNSET, np.arange(ntot), pc)
    res["copy_ok"] = bool(np.array_equal(L1, L0[src]) and np.array_equal(Li1, Li0[src]) and np.array_equal(sums1, sums0[src]))
    # the copied particles carry a complete state: a model-based proposal from them works without a replay
    prop = eng.batch(N, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    prop.set_params(theta)
    prop.mbp(b, None, seed=5)
    res["mbp_finite"] = bool(np.all(np.isfinite(prop.obs_loglik())))
    out[rank] = res
    cm.close()


def _run(world, name):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ctx = mp.get_context("spawn")
    out = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        idfile = os.path.join(td, "nccl_id")
        procs = [ctx.Process(target=_worker, args=(r, world, name, idfile, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(600)
            assert p.exitcode == 0
    return dict(out)


@pytest.mark.parametrize("name", ["EX_C4", "EX_C2"])
def test_multi_gpu_collectives_equal_one_gpu(name):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    one = _run(1, name)[0]
    worlds = [2] + ([4] if ngpu >= 4 else []) + ([8] if ngpu >= 8 else [])
    for world in worlds:
        res = _run(world, name)
        N = 96 // world
        for r in range(world):
            assert res[r]["gather"] == [float(k + 1000 * q) for q in range(world) for k in range(N)]
            assert res[r]["back"] == one["back"], "ranks disagree with the one-GPU ancestry"
            assert res[r]["ll"] == one["ll"], "sharded filter changed the likelihood"
            assert res[r]["copy_ok"] and res[r]["mbp_finite"]
        assert res[0]["sample"] == one["sample"], "particle_sample differs from the one-GPU trajectory"
        assert sum(res[r]["swap_bytes"][0] for r in range(world)) == sum(res[r]["swap_bytes"][1] for r in range(world)) > 0


def test_cpp_filter_two_ranks(tmp_path):
    """The C++ host layer over two GPUs (include/beep_b200.hpp: beep::Comm + beep::particle_filter, host/pmcmc_filter.cc,
    one process per GPU): the same log marginal likelihood, to the last bit, as all particles on one GPU."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from beep_b200.bpk import write_bpk
    from beep_b200.model import ModelDesc
    exe = os.path.join(ROOT, "host", "pmcmc_filter")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    m, d = load_golden("EX_C4")
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    out = ModelDesc(arrays).to_prefixed()
    out["theta"] = np.ascontiguousarray(d["theta"][0], dtype=np.float64)
    path = str(tmp_path / "model.bpk")
    write_bpk(path, out)
    one = subprocess.run([exe, path, "64", "1", "4"], capture_output=True, text=True, timeout=300, env=dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0"))
    assert one.returncode == 0, one.stderr
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([exe, path, "64", "1", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=dict(os.environ, WORLD_SIZE="2", RANK=str(r), LOCAL_RANK=str(r), BEEP_NCCL_ID_FILE=idfile)) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    ll1 = one.stdout.split("log-likelihood=")[1].split()[0]
    ll2 = outs[0][0].split("log-likelihood=")[1].split()[0]
    assert "ranks=2" in outs[0][0] and ll1 == ll2, (one.stdout, outs[0][0])
