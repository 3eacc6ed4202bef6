"""The send / receive lists the C-ABI library derives for Mpi::copy_particles (mpi.cc:81-165) and
Mpi::end_sec_swap (mpi.cc:215-327), replayed here for EVERY rank of a 2-, 4- and 8-rank job on plain
arrays: messages of a pair of ranks are matched in posting order (NCCL's rule), copies within a rank are
out of place, staged arrivals land last -- the outcome must be the global permutation the decision
vector describes, and end_sec_swap must ship each distinct remote ancestor once per destination rank.
No GPU needed: this is the host logic of beep_b200/csrc/comm.cuh."""
import numpy as np
import pytest

UNSET = 0xFFFFFFFF


def _replay(kind, vec, N, world):
    from beep_b200 import engine as E
    ntot = N * world
    state = np.arange(ntot) * 10 + 7            # the "content" of every global particle
    plans = [E.move_plan(kind, vec, N, r) for r in range(world)]
    # post: per ordered pair (src, dst) a FIFO of messages, in each side's own posting order
    sent = {}
    for r, (_, out, _) in enumerate(plans):
        for peer, slot in out:
            sent.setdefault((r, peer), []).append(state[r * N + slot])
    new = state.copy()
    nmsg = 0
    for r, (local, _, ins) in enumerate(plans):
        taken = {}
        staged = []
        for peer, slots in ins:
            k = taken.get(peer, 0)
            q = sent.get((peer, r), [])
            assert k < len(q), f"rank {r} expects a message from {peer} that was never sent"
            staged.append((q[k], slots))
            taken[peer] = k + 1
            nmsg += 1
        for peer, k in taken.items():
            assert k == len(sent[(peer, r)]), "unmatched sends"
        old = state[r * N:(r + 1) * N].copy()
        for k in range(N):
            if local[k] != UNSET:
                new[r * N + k] = old[local[k]]
        for val, slots in staged:
            for s in slots:
                new[r * N + s] = val
    for (src, dst), q in sent.items():
        assert sum(1 for peer, _ in plans[dst][2] if peer == src) == len(q)
    return state, new, nmsg


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", [0, 1])
def test_plans_realise_the_global_permutation(kind, world):
    rng = np.random.default_rng(100 * kind + world)
    N = 6
    ntot = N * world
    for trial in range(40):
        if kind == 0:
            vec = np.full(ntot, UNSET, dtype=np.uint32)
            idx = rng.choice(ntot, size=rng.integers(0, ntot + 1), replace=False)
            vec[idx] = rng.integers(0, ntot, size=idx.size)
        else:
            vec = rng.integers(0, ntot, size=ntot).astype(np.uint32)       # ancestors: many particles share few
            if trial % 3 == 0:
                vec = rng.choice(rng.integers(0, ntot, size=3), size=ntot).astype(np.uint32)
        state, new, nmsg = _replay(kind, vec, N, world)
        want = state.copy()
        for i in range(ntot):
            if vec[i] != UNSET:
                want[i] = state[vec[i]]
        assert np.array_equal(new, want)
        if kind == 1:
            # one message per (distinct remote ancestor, destination rank) -- mpi.cc:243-258,270-287
            pairs = {(int(vec[i]), i // N) for i in range(ntot) if vec[i] // N != i // N}
            assert nmsg == len(pairs)


def test_keep_and_self_entries_move_nothing():
    from beep_b200 import engine as E
    N, world = 4, 2
    vec = np.array([UNSET, 1, UNSET, 3, 4, UNSET, 6, UNSET], dtype=np.uint32)
    for r in range(world):
        local, out, ins = E.move_plan(0, vec, N, r)
        assert np.all(local == UNSET) and not out and not ins
