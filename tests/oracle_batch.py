"""A CPU stand-in for beep_b200.engine.Batch built on the oracle -- TEST INFRASTRUCTURE.
It lets the multi-rank host logic (beep_b200/shard.py) run under gloo without a GPU."""
import numpy as np

from oracle_lib import OracleState


class OracleEngine:
    def __init__(self, model):
        self.model = model
        self._st = OracleState(model, stochastic=1, keyed=True, imap_from_pop=True)
        self.nsection = self._st.nsection

    def section_range(self, s):
        return self._st.section_range(s)


class OracleBatch:
    def __init__(self, engine, P):
        self.engine, self.model, self.P = engine, engine.model, P
        self.st = [OracleState(engine.model, stochastic=1, keyed=True, imap_from_pop=True) for _ in range(P)]
        self.ids = np.arange(P, dtype=np.uint64)
        self.cur_t = 0

    def set_params(self, theta):
        th = np.asarray(theta, dtype=np.float64)
        if th.ndim == 1:
            th = np.broadcast_to(th, (self.P, th.size))
        self.theta = np.array(th, dtype=np.float64)
        for k in range(self.P):
            self.st[k].set_param(th[k])

    def simulate(self, ti=0, tf=None, seed=0, stream_ids=None):
        tf = self.model.T if tf is None else tf
        if stream_ids is not None:
            self.ids = np.asarray(stream_ids, dtype=np.uint64).copy()
        for k in range(self.P):
            self.st[k].simulate(ti, tf, seed=seed, stream_id=int(self.ids[k]))
        self.cur_t = tf

    def obs_loglik(self, section=-1):
        return np.array([s.obs_loglik(section) for s in self.st])

    def pop_all(self):
        t = min(self.cur_t, self.model.T - 1)
        return np.stack([s.pop[t].copy() for s in self.st])

    def set_pop_all(self, pop, cur_step):
        for k in range(self.P):
            self.st[k].pop[cur_step] = pop[k]
        self.cur_t = cur_step

    def permute(self, src_of, what=0):
        src = [k if s == 0xFFFFFFFF else int(s) for k, s in enumerate(src_of)]
        if what == 0:                                   # PERMUTE_FULL: struct Particle = (paramval, transnum)
            snap = [self.export_particle(k) for k in range(self.P)]
            for k, s in enumerate(src):
                if s != k:
                    self.import_particle(k, *snap[s])
            return
        pops = self.pop_all()
        self.set_pop_all(pops[src], self.cur_t)

    def export_particle(self, slot):
        return self.theta[slot].copy(), self.st[slot].transnum.copy()

    def import_particle(self, slot, theta, transnum):
        self.theta[slot] = theta
        self.st[slot].set_param(theta)
        self.st[slot].init_from_particle(np.ascontiguousarray(transnum, dtype=np.float64))

    def replay(self):
        pass                                            # init_from_particle already rebuilt the state
