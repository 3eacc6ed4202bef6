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
