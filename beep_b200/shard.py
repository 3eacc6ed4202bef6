"""Particle sharding across the GPUs of one box -- the replacement of the reference's MPI ranks.

The reference gives each rank ``N = Ntot/ncore`` particles (``inputs.cc:1797-1832``), gathers
per-particle scalars to rank 0, decides there with the global RNG, broadcasts the decision and moves
particles pairwise (``mpi.cc:81-165,215-327``).  Here every rank all-gathers the scalars (8 bytes per
particle) and computes the IDENTICAL decision from a shared counter-based key, so no root and no
broadcast exist; only the states of ancestors that live on another rank move.

The simulate / MBP / likelihood path itself has no collective.  Everything here is host logic over
``torch.distributed`` (backend nccl on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def partition(ntot, world):
    """Global particle i lives on rank i // (ntot // world) -- the reference's core = i/N (mpi.cc:98-101)."""
    if ntot % world != 0:
        raise ValueError(f"the number of particles ({ntot}) must be a multiple of the number of ranks ({world}), as in the reference (inputs.cc:1810)")
    return ntot // world


def owner(i, ntot, world):
    return i // partition(ntot, world)


def local_range(rank, ntot, world):
    n = partition(ntot, world)
    return rank * n, (rank + 1) * n


def allgather(local, dist=None, device=None):
    """All-gather a per-particle array (first axis = local particles) from every rank, in rank order."""
    local = np.ascontiguousarray(local)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local.copy()
    import torch
    t = torch.from_numpy(local)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy() for o in out], axis=0)


def ef_cutoff_copy(EF_all):
    """ABCMBP::EF_cutoff (abcmbp.cc:135-217), the decision part: sort the error functions, the cut-off
    is the mean of the two middle values (:149-150), the j-th best particle is kept and the
    (j + Ntot/2)-th is overwritten by a copy of it (:158-163).
    Returns (EFcut, EFmin, EFmax, partcopy) with partcopy[i] = source particle or 0xFFFFFFFF (keep)."""
    EF_all = np.asarray(EF_all, dtype=np.float64)
    n = EF_all.size
    order = np.argsort(EF_all, kind="stable")
    half = n // 2
    EFcut = 0.5 * (EF_all[order[half - 1]] + EF_all[order[half]])
    EFmin, EFmax = EF_all[order[int(0.025 * n)]], EF_all[order[int(0.975 * n)]]
    partcopy = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    for j in range(half):
        partcopy[order[j + half]] = order[j]
    return float(EFcut), float(EFmin), float(EFmax), partcopy


class ShardedFilter:
    """Bootstrap particle filter of PMCMC::obs_prob (pmcmc.cc:180-268) over particle shards.

    Every rank holds ``batch`` (its N particles, all with the same parameters).  Per section:
    simulate, per-particle section likelihood, all-gather of L, replicated bootstrap, and the
    ancestor states are fetched from the all-gathered section state.  Returns the log marginal
    likelihood estimate, identical on every rank.
    """

    def __init__(self, batch, rank=0, world=1, dist=None, device=None):
        self.batch = batch
        self.rank, self.world = rank, world
        self.dist = dist
        self.device = device
        self.ancestry = []

    def run(self, theta, seed, invT=1.0):
        from . import engine as E
        b = self.batch
        eng = b.engine
        N = b.P
        ntot = N * self.world
        lo = self.rank * N
        b.set_params(theta)
        ids = np.arange(lo, lo + N, dtype=np.uint64)
        loglik = 0.0
        self.ancestry = []
        for s in range(eng.nsection):
            ti, tf = eng.section_range(s)
            if s > 0:
                back = self.ancestry[-1]
                mine = back[lo:lo + N]
                if self.world == 1:
                    b.permute(mine, E.PERMUTE_SECTION_STATE)
                else:
                    allpop = allgather(b.pop_all(), self.dist, self.device)      # [Ntot][A][C][D]
                    b.set_pop_all(allpop[mine], ti)
            b.simulate(ti, tf, seed=seed, stream_ids=ids if s == 0 else None)
            L = b.obs_loglik(section=s)
            L_all = allgather(L, self.dist, self.device)
            assert L_all.size == ntot
            inc, back = E.bootstrap(L_all, invT, seed, s)
            loglik += inc
            self.ancestry.append(back)
        return loglik
