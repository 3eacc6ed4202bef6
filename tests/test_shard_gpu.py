"""Particle shards on two real GPUs over NCCL (skipped on a one-GPU box).

The particle filter of PMCMC::obs_prob (pmcmc.cc:180-268) runs (a) on one GPU with 2N particles and
(b) on two GPUs with N particles each -- section likelihoods all-gathered over NCCL, the bootstrap
decision replicated, ancestor states fetched from the all-gathered section state.  Same particles,
same keyed streams: log-likelihood and ancestry must agree bit for bit (the reference's equivalent
is a run with 1 or 2 MPI ranks and the same Ntot, mpi.cc:215-327).
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _filter(rank, world, N, name, seed, port, out):
    import torch
    import torch.distributed as dist
    from beep_b200 import engine as E
    from beep_b200.shard import ShardedFilter
    m, d = load_golden(name)
    torch.cuda.set_device(rank)
    d_ = None
    dev = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        d_ = dist
        dev = torch.device("cuda", rank)
    eng = E.Engine(m, rank)
    b = eng.batch(N, 0)
    f = ShardedFilter(b, rank, world, d_, dev)
    ll = f.run(d["theta"][0], seed)
    out[rank] = (ll, [a.tolist() for a in f.ancestry])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["EX_A1", "EX_C4", "EX_C2"])
def test_two_gpu_filter_equals_one_gpu(name):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    N = 64
    mgr = mp.Manager()
    ctx = mp.get_context("spawn")
    out1 = mgr.dict()
    p1 = ctx.Process(target=_filter, args=(0, 1, 2 * N, name, 11, 0, out1))
    p1.start()
    p1.join(300)
    assert p1.exitcode == 0
    out2 = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_filter, args=(r, 2, N, name, 11, port, out2)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out2[0] == out2[1], "ranks disagree on the replicated decision"
    assert out2[0][0] == out1[0][0], "sharded filter changed the likelihood"
    assert out2[0][1] == out1[0][1], "sharded filter changed the ancestry"
    assert np.isfinite(out1[0][0])


def _mover(rank, world, N, name, port, src_of_all, out):
    import torch
    import torch.distributed as dist
    from beep_b200 import engine as E
    from beep_b200.shard import copy_particles
    m, d = load_golden(name)
    torch.cuda.set_device(rank)
    d_, dev = None, None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        d_, dev = dist, torch.device("cuda", rank)
    from beep_b200.model import ModelDesc
    arrays = dict(m.arrays)
    arrays["stochastic"] = 1
    eng = E.Engine(ModelDesc(arrays), rank)
    b = eng.batch(N, E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN)
    th = np.repeat(d["theta"][:1], N, axis=0) * (1 + 0.001 * (rank * N + np.arange(N))[:, None])
    b.set_params(th)
    b.simulate(seed=5, stream_ids=np.arange(rank * N, (rank + 1) * N, dtype=np.uint64))
    copy_particles(b, src_of_all, rank, world, d_, dev)
    L = b.obs_loglik()
    Li = b.latent_loglik()          # depends on the derived tables of the moved parameters
    out[rank] = [(b.export_particle(k)[0].tolist(), b.export_particle(k)[1].ravel().tolist(), float(L[k]), float(Li[k])) for k in range(N)]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_copy_particles_across_gpus():
    """Mpi::copy_particles (mpi.cc:81-165) over two GPUs: particles travel as struct Particle point to point
    over NCCL and are rebuilt by replay; equal to the same permutation on one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from beep_b200.shard import UNSET
    N = 3
    src = np.array([UNSET, 4, 0, 1, UNSET, 3], dtype=np.uint32)
    mgr = mp.Manager()
    ctx = mp.get_context("spawn")
    one = mgr.dict()
    p1 = ctx.Process(target=_mover, args=(0, 1, 2 * N, "EX_C4", 0, src, one))
    p1.start()
    p1.join(300)
    assert p1.exitcode == 0
    two = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_mover, args=(r, 2, N, "EX_C4", port, src, two)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    hung = [p for p in procs if p.exitcode is None]
    for p in hung:
        p.kill()
    assert not hung and all(p.exitcode == 0 for p in procs)
    a, b = two[0] + two[1], one[0]
    for x, y in zip(a, b):
        assert x[0] == y[0] and x[1] == y[1]
        assert abs(x[2] - y[2]) <= 1e-10 * abs(y[2])
        assert abs(x[3] - y[3]) <= 1e-10 * abs(y[3])
