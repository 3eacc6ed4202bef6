"""Multi-rank host logic (beep_b200/shard.py) under gloo, world_size 2, on CPU.

The particle filter of PMCMC::obs_prob (pmcmc.cc:180-268) is run (a) on one rank with 2N particles
and (b) on two ranks with N particles each, all-gathering the section likelihoods and the ancestor
states; both must give the same log-likelihood and the same ancestry, bit for bit.  Compute comes
from the CPU oracle (tests/oracle_batch.py); the bootstrap is the C-ABI's host routine.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _filter(rank, world, N, name, seed, port, out):
    import torch.distributed as dist
    from beep_b200.shard import ShardedFilter
    from oracle_batch import OracleBatch, OracleEngine
    m, d = load_golden(name)
    d_ = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        d_ = dist
    eng = OracleEngine(m)
    b = OracleBatch(eng, N)
    f = ShardedFilter(b, rank, world, d_)
    ll = f.run(d["theta"][0], seed)
    out[rank] = (ll, [a.tolist() for a in f.ancestry])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["EX_A1", "EX_C4"])
def test_two_rank_filter_equals_single_rank(name):
    import torch.multiprocessing as mp
    N = 3
    mgr = mp.Manager()
    out1 = mgr.dict()
    _filter(0, 1, 2 * N, name, 11, 0, out1)
    out2 = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
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


def test_partition_matches_reference_layout():
    from beep_b200 import shard
    assert shard.partition(16384, 8) == 2048
    assert [shard.owner(i, 16, 4) for i in (0, 3, 4, 15)] == [0, 0, 1, 3]
    assert shard.local_range(2, 16, 4) == (8, 12)
    with pytest.raises(ValueError):
        shard.partition(10, 4)


def test_ef_cutoff_copy():
    """ABCMBP::EF_cutoff (abcmbp.cc:135-217): best half overwrites worst half"""
    from beep_b200 import shard
    EF = np.array([5.0, 1.0, 9.0, 3.0, 7.0, 2.0, 8.0, 4.0])
    cut, lo, hi, pc = shard.ef_cutoff_copy(EF)
    assert cut == 0.5 * (4.0 + 5.0)
    keep = EF < cut
    assert np.all(pc[keep] == 0xFFFFFFFF)
    src = pc[~keep]
    assert sorted(src.tolist()) == sorted(np.nonzero(keep)[0].tolist())
    # j-th best feeds the (j + N/2)-th best
    order = np.argsort(EF)
    for j in range(4):
        assert pc[order[j + 4]] == order[j]


def test_pas_bootstrap_decision():
    """PAS::bootstrap (pas.cc:127-209): the replicated decision against a literal restatement of the
    reference's loops fed with the same uniforms, plus its invariants."""
    from beep_b200.shard import pas_bootstrap, _uniforms, UNSET
    rng = np.random.default_rng(3)
    nrun, npart = 3, 40
    EF = rng.gamma(5.0, 20.0, nrun * npart)
    invT, quench = 0.01, 0.5
    DinvT, EFmin, EFmax, pc = pas_bootstrap(EF, invT, quench, nrun, npart, seed=9, gen=4)
    srt = np.sort(EF)
    n = EF.size
    assert DinvT == 0.25 * quench / (srt[int(0.5 * n)] - srt[int(0.25 * n)])
    assert (EFmin, EFmax) == (srt[int(0.25 * n)], srt[int(0.75 * n)])
    u = _uniforms(9, 4, 0x9A5, n)
    muav = EF.sum() / n
    for ru in range(nrun):
        wsum, s = [], 0.0
        for j in range(npart):
            s += np.exp(-0.5 * DinvT * (EF[ru * npart + j] - muav))
            wsum.append(s)
        num = [0] * npart
        for j in range(npart):
            z = u[ru * npart + j] * s
            sel = 0
            while sel < npart - 1 and z > wsum[sel]:
                sel += 1
            num[sel] += 1
        lst = [ru * npart + j for j in range(npart) for _ in range(max(num[j] - 1, 0))]
        for j in range(npart):
            if num[j] > 0:
                assert pc[ru * npart + j] == UNSET
            else:
                assert pc[ru * npart + j] == lst.pop()
                assert ru * npart <= pc[ru * npart + j] < (ru + 1) * npart       # copies never cross runs
    # the inverse-temperature step is capped by invT_final (pas.cc:156-158)
    assert pas_bootstrap(EF, invT, quench, nrun, npart, 9, 4, invT_final=invT + 1e-6)[0] == pytest.approx(1e-6)
    # replicated: the same key gives the same decision, another generation another one
    assert np.array_equal(pc, pas_bootstrap(EF, invT, quench, nrun, npart, 9, 4)[3])
    assert not np.array_equal(pc, pas_bootstrap(EF, invT, quench, nrun, npart, 9, 5)[3])


def test_mc3_swap_decision():
    """MC3::swap_states (mc3.cc:153-217): swaps stay inside a run, hot/cold pairs with the colder chain at the
    higher error always swap (al >= 1), equal temperatures always swap, counts add up."""
    from beep_b200.shard import mc3_swap
    nrun, nchain = 2, 6
    run = np.repeat(np.arange(nrun), nchain)
    invT = np.tile(np.linspace(1.0, 0.1, nchain), nrun)
    rng = np.random.default_rng(4)
    EF = rng.uniform(50, 150, run.size)
    order, ntr, nac = mc3_swap(run, invT, EF, seed=2, sweep=0)
    assert sorted(order.tolist()) == list(range(run.size))
    for r in range(nrun):
        seg = order[r * nchain:(r + 1) * nchain]
        assert np.all((seg >= r * nchain) & (seg < (r + 1) * nchain))
    assert ntr[nchain - 1] == 0 and ntr[0] == 5 and np.all(nac <= ntr)         # no pair across the run boundary
    # a colder chain (larger invT) sitting on the larger EF: al = exp(+) >= 1, the first pass swaps for sure
    o2, t2, a2 = mc3_swap([0, 0], [1.0, 0.5], [100.0, 10.0], seed=1, sweep=0)
    assert a2[0] >= 1
    # equal temperatures: al = 1, every proposal accepted -> five swaps = one net swap
    o3, t3, a3 = mc3_swap([0, 0], [0.7, 0.7], [3.0, 8.0], seed=1, sweep=0)
    assert a3[0] == 5 and o3.tolist() == [1, 0]
    assert np.array_equal(order, mc3_swap(run, invT, EF, 2, 0)[0])


def _mover(rank, world, N, name, port, src_of_all, out):
    import torch.distributed as dist
    from beep_b200.shard import copy_particles
    from oracle_batch import OracleBatch, OracleEngine
    m, d = load_golden(name)
    d_ = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        d_ = dist
    b = OracleBatch(OracleEngine(m), N)
    th = np.repeat(d["theta"][:1], N, axis=0) * (1 + 0.001 * (rank * N + np.arange(N))[:, None])
    b.set_params(th)
    b.simulate(seed=5, stream_ids=np.arange(rank * N, (rank + 1) * N, dtype=np.uint64))
    copy_particles(b, src_of_all, rank, world, d_)
    out[rank] = [(b.export_particle(k)[0].tolist(), b.export_particle(k)[1].ravel().tolist(), float(b.obs_loglik()[k])) for k in range(N)]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_copy_particles_across_ranks():
    """Mpi::copy_particles (mpi.cc:81-165): the same global permutation on one rank with 2N particles and on
    two ranks with N each (copies inside a rank, across ranks in both directions, chains of copies)."""
    import torch.multiprocessing as mp
    from beep_b200.shard import UNSET
    N = 3
    src = np.array([UNSET, 4, 0, 1, UNSET, 3], dtype=np.uint32)     # 1<-4 and 3<-1 cross ranks, 2<-0 and 5<-3 stay
    mgr = mp.Manager()
    one = mgr.dict()
    _mover(0, 1, 2 * N, "EX_C4", 0, src, one)
    two = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_mover, args=(r, 2, N, "EX_C4", port, src, two)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert two[0] + two[1] == one[0]
    # out of place: particle 3 received the ORIGINAL particle 1, not the copy of 4 that landed in slot 1
    assert one[0][3][0] != one[0][1][0] and one[0][1][0] == one[0][4][0]
