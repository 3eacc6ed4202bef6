"""Particle sharding across the GPUs of one box -- the replacement of the reference's MPI ranks.

The reference gives each rank ``N = Ntot/ncore`` particles (``inputs.cc:1797-1832``), gathers
per-particle scalars to rank 0, decides there with the global RNG, broadcasts the decision and moves
particles pairwise (``mpi.cc:81-165,215-327``).  Here every rank all-gathers the scalars (8 bytes per
particle) and computes the IDENTICAL decision from a shared counter-based key, so no root and no
broadcast exist; only the states of ancestors that live on another rank move.

The simulate / MBP / likelihood path itself has no collective.  The decisions themselves (bootstrap, cut-off, PAS resampling, MC3 swaps) are host
routines of the C-ABI library (C++); this module is the ``torch.distributed`` plumbing around them (backend nccl on
the GPU box, gloo in the CPU tests).
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
    """ABCMBP::EF_cutoff (abcmbp.cc:135-217), the decision part (beep_ef_cutoff): sort the error functions,
    the cut-off is the mean of the two middle values (:149-150), the j-th best particle is kept and the
    (j + Ntot/2)-th is overwritten by a copy of it (:158-163).
    Returns (EFcut, EFmin, EFmax, partcopy) with partcopy[i] = source particle or 0xFFFFFFFF (keep)."""
    import ctypes
    from . import engine as E
    EF_all = np.ascontiguousarray(EF_all, dtype=np.float64)
    pc = np.zeros(EF_all.size, dtype=np.uint32)
    cut, lo, hi = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    E._check(E.lib().beep_ef_cutoff(EF_all.ctypes.data, EF_all.size, ctypes.byref(cut), ctypes.byref(lo), ctypes.byref(hi), pc.ctypes.data))
    return cut.value, lo.value, hi.value, pc


class ShardedFilter:
    """Bootstrap particle filter of PMCMC::obs_prob (pmcmc.cc:180-268) over particle shards.

    Every rank holds ``batch`` (its N particles, all with the same parameters).  Per section:
    simulate, per-particle section likelihood, all-gather of L, replicated bootstrap, and the
    ancestor states are fetched from the all-gathered section state.  Returns the log marginal
    likelihood estimate, identical on every rank.
    """

    def __init__(self, batch, rank=0, world=1, dist=None, device=None, comm=None):
        self.batch = batch
        self.rank, self.world = rank, world
        self.dist = dist
        self.device = device
        self.comm = comm              # engine.Comm: the library's own NCCL collectives (device to device)
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
                elif self.comm is not None:
                    # Mpi::end_sec_swap (mpi.cc:215-327): only the distinct ancestors owned by another rank travel, HBM to HBM
                    self.comm.end_sec_swap(b, back)
                else:
                    # CPU test path (gloo, oracle batches): the ancestors' states from the all-gathered section state
                    allpop = allgather(b.pop_all(), self.dist, self.device)      # [Ntot][A][C][D]
                    b.set_pop_all(allpop[mine], ti)
            b.simulate(ti, tf, seed=seed, stream_ids=ids if s == 0 else None)
            L = b.obs_loglik(section=s)
            L_all = self.comm.allgather(L) if (self.comm is not None and self.world > 1) else allgather(L, self.dist, self.device)
            assert L_all.size == ntot
            inc, back = E.bootstrap(L_all, invT, seed, s)
            loglik += inc
            self.ancestry.append(back)
        return loglik


# ---------------------------------------------------------------------------
# a12: the resampling decisions of PAS and MC3, and the particle mover
# ---------------------------------------------------------------------------
UNSET = 0xFFFFFFFF


def _uniforms(seed, key, tag, n):
    """the uniforms a decision consumes (test hook of the library)"""
    from . import engine as E
    out = np.zeros(n)
    E._check(E.lib().beep_test_uniforms(int(seed), int(key), int(tag), n, out.ctypes.data))
    return out


def pas_bootstrap(EF_all, invT, quench_factor, nrun, npart, seed, gen, invT_final=None):
    """PAS::bootstrap (pas.cc:127-209), the decision part (beep_pas_bootstrap), on the all-gathered error
    functions: the step in inverse temperature is 0.25 quench_factor / (median - lower quartile of EF)
    (:147-154), limited by invT_final (:156-158); per run the particles are resampled in proportion to
    exp(-DinvT (EF - mean)/2) (:160-176); a particle drawn k > 1 times is copied k-1 times over particles
    drawn 0 times, handed out from the END of the list as the reference does (:178-193).
    Returns (DinvT, EFmin, EFmax, partcopy) with partcopy[i] = source particle or UNSET (keep)."""
    import ctypes
    from . import engine as E
    EF_all = np.ascontiguousarray(EF_all, dtype=np.float64)
    if EF_all.size != nrun * npart:
        raise ValueError("EF_all must hold nrun*npart values")
    pc = np.zeros(EF_all.size, dtype=np.uint32)
    d, lo, hi = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    E._check(E.lib().beep_pas_bootstrap(EF_all.ctypes.data, nrun, npart, float(invT), float(quench_factor), -1.0 if invT_final is None else float(invT_final),
                                        int(seed), int(gen), ctypes.byref(d), ctypes.byref(lo), ctypes.byref(hi), pc.ctypes.data))
    return d.value, lo.value, hi.value, pc


def mc3_swap(run_all, invT_all, EF_all, seed, sweep):
    """MC3::swap_states (mc3.cc:153-217), the decision part (beep_mc3_swap): five passes over neighbouring
    chains of the same run, swap accepted with probability exp((invT_i - invT_{i+1}) (EF_i - EF_{i+1}) / 2)
    (:178-190).  Returns (order, ntr_inc, nac_inc): chain position i continues with the particle that was
    at order[i] (the argument of Mpi::copy_particles, :207) and the proposal / acceptance counts to add."""
    from . import engine as E
    run = np.ascontiguousarray(run_all, dtype=np.uint32)
    invT = np.ascontiguousarray(invT_all, dtype=np.float64)
    EF = np.ascontiguousarray(EF_all, dtype=np.float64)
    n = EF.size
    order, ntr, nac = (np.zeros(n, dtype=np.uint32) for _ in range(3))
    E._check(E.lib().beep_mc3_swap(run.ctypes.data, invT.ctypes.data, EF.ctypes.data, n, int(seed), int(sweep), order.ctypes.data, ntr.ctypes.data, nac.ctypes.data))
    return order, ntr, nac


def copy_particles(batch, src_of_all, rank=0, world=1, dist=None, device=None):
    """Mpi::copy_particles (mpi.cc:81-165) over particle shards: global particle i becomes a copy of global
    particle src_of_all[i] (UNSET = keep).  Copies within a rank are one permutation on the device
    (beep_permute, out of place); a particle whose source lives on another rank travels as struct Particle
    -- (paramval, transnum) -- point to point, and its state is rebuilt by replay
    (State::initialise_from_particle, state.cc:294-322).  All sources are read before any destination is
    written, as the reference's send-then-overwrite order guarantees."""
    src_of_all = np.asarray(src_of_all, dtype=np.uint32)
    N = batch.P
    lo = rank * N
    mine = src_of_all[lo:lo + N]
    local = np.full(N, UNSET, dtype=np.uint32)
    incoming = []                                        # (local slot, source rank, source slot)
    for k, s in enumerate(mine):
        if s == UNSET or int(s) == lo + k:
            continue
        if lo <= int(s) < lo + N:
            local[k] = int(s) - lo
        else:
            incoming.append((k, int(s) // N, int(s) % N))
    outgoing = []                                        # (destination rank, local source slot, global destination)
    if world > 1:
        for i, s in enumerate(src_of_all):
            if s != UNSET and lo <= int(s) < lo + N and not (lo <= i < lo + N):
                outgoing.append((i // N, int(s) - lo, i))
    received = []
    if world > 1 and (incoming or outgoing):
        import torch
        dev = device if device is not None else torch.device("cpu")
        ops, keep = [], []
        for dst_rank, slot, gdst in outgoing:                # read every source first
            th, tn = batch.export_particle(slot)
            t = torch.from_numpy(np.concatenate([th.ravel(), tn.ravel()])).to(dev)
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, dst_rank, tag=gdst))
        m = batch.model
        nth, ntn = int(m.nparam), int(np.prod(m.traj_shape))
        for k, src_rank, _ in incoming:
            t = torch.empty(nth + ntn, dtype=torch.float64, device=dev)
            ops.append(dist.P2POp(dist.irecv, t, src_rank, tag=lo + k))
            received.append((k, t))
        # one group: NCCL matches sends and receives of a pair in posting order (both sides list them by
        # global destination) and would deadlock on send-before-receive outside a group
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    if np.any(local != UNSET):
        batch.permute(local)
    if received:
        m = batch.model
        nth = int(m.nparam)
        for k, t in received:
            v = t.cpu().numpy()
            batch.import_particle(k, v[:nth], v[nth:].reshape(m.traj_shape))
        batch.replay()
