#!/usr/bin/env python
"""Times the auxiliary kernels of the path on the synthetic population (one B200): K5 latent-process likelihood
(State::likelihood, state.cc:812-826), the parameter-proposal replay (Mbp::parameter_proposal, mbp.cc:460-491:
set_param + transition means on fixed events + likelihood), K6 particle copies (Mpi::copy_particles within a GPU,
Mbp::swap_initial_propose_state).  usage: python profiles/time_aux.py [areas] [particles]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beep_b200 import engine as E
from beep_b200.synth import make_model, jitter_params
A = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m, theta = make_model(A=A, G=20, T=365)
eng = E.Engine(m)
flags = E.KEEP_TRANSNUM | E.KEEP_TRANSMEAN
a, b = eng.batch(P, flags), eng.batch(P, flags)
th = jitter_params(theta, P, rel=0.05, seed=3)
for x in (a, b):
    x.set_params(th)
    x.simulate(seed=5)
    x.sync()
cells = P * m.T * m.A * m.J * m.D


def timed(fn, reps=5):
    fn(); a.sync(); b.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    a.sync(); b.sync()
    return (time.perf_counter() - t0) / reps


t = timed(lambda: a.latent_loglik())
print(f"K5 latent log-likelihood: {t*1e3:.2f} ms, {cells/t/1e9:.1f} G cells/s, {cells*12/t/1e9:.0f} GB/s of the 12 B/cell it reads")
t = timed(lambda: (a.set_params(th * 1.001), a.replay(), a.latent_loglik()))
print(f"parameter proposal (set_param + replay on fixed events + likelihood): {t*1e3:.2f} ms, {cells/t/1e9:.2f} G cells/s")
src = np.roll(np.arange(P, dtype=np.uint32), 1)
t = timed(lambda: a.permute(src, E.PERMUTE_FULL))
slab = cells * 12 + P * m.A * m.C * m.D * 8
print(f"K6 permute (every particle replaced: gather into a compact scratch, write into place): {t*1e3:.2f} ms, {2*slab/t/1e9:.0f} GB/s read+write of the slabs once each, {a.engine.launch_count} launches so far")
n0 = a.engine.launch_count
rng = np.random.default_rng(1)
src = np.sort(rng.integers(0, P, size=P)).astype(np.uint32)          # a multinomial resampling: about 1/e of the slots receive another particle
keep = np.zeros(P, dtype=bool); keep[np.unique(src)] = True
srcm = np.where(keep, np.arange(P), 0).astype(np.uint32)
dup = [s_ for s_ in src if True]
free = [i for i in range(P) if not keep[i]]
extra = []
seen = set()
for s_ in src:
    if int(s_) in seen: extra.append(int(s_))
    seen.add(int(s_))
for i, s_ in zip(free, extra): srcm[i] = s_
nmoved = int((srcm != np.arange(P)).sum())
t = timed(lambda: a.permute(srcm, E.PERMUTE_FULL), reps=1)
print(f"K6 permute after a multinomial resampling ({nmoved} of {P} slots receive another particle): {t*1e3:.2f} ms, {(a.engine.launch_count - n0) // 2} launches per call")
acc = np.ones(P, dtype=np.uint8)
t = timed(lambda: a.accept_from(b, acc))
print(f"K6 accept (swap_initial_propose_state for every chain): {t*1e3:.2f} ms, {2*slab/t/1e9:.0f} GB/s read+write")
