"""ctypes binding of oracle/liboracle.so -- the CPU restatement used as the
parity checker.  TEST INFRASTRUCTURE: only tests/, smoke() and the bench's
cpu_baseline leg import this module."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

TRANSRATE, DISC_SPLINE, SUS, AREAFACTOR, BETA, NTIME, POP, TRANSNUM, TRANSMEAN, IMAP, IDIAG, PARAMVAL = range(12)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        L = ctypes.CDLL(path)
        L.oracle_state_create.restype = ctypes.c_void_p
        L.oracle_state_create.argtypes = [ctypes.c_void_p]
        L.oracle_state_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_set_mode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_ptr.restype = ctypes.POINTER(ctypes.c_double)
        L.oracle_ptr.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_nsection.argtypes = [ctypes.c_void_p]
        L.oracle_section_range.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.oracle_ndraw.restype = ctypes.c_long
        L.oracle_ndraw.argtypes = [ctypes.c_void_p]
        L.oracle_set_param.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_simulate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64]
        L.oracle_init_from_particle.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_recompute_transmean.argtypes = [ctypes.c_void_p]
        L.oracle_copy_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_mbp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.oracle_obs_values.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_obs_loglik.restype = ctypes.c_double
        L.oracle_obs_loglik.argtypes = [ctypes.c_void_p]
        L.oracle_obs_loglik_section.restype = ctypes.c_double
        L.oracle_obs_loglik_section.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_EF_datatable.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_latent_loglik.restype = ctypes.c_double
        L.oracle_latent_loglik.argtypes = [ctypes.c_void_p]
        L.oracle_bootstrap.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
        L.oracle_poisson_many.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_binomial_many.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_philox.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_poisson_inv_many.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_exp_many.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_log_many.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_mvn_kernel_sums.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_graph_npoints.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_graph_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_likelihood_approx.restype = ctypes.c_double
        L.oracle_likelihood_approx.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def mvn_kernel_sums(var, inv_M, size, param_prop, param_last, w_last):
    """SUM_j w_last[j] exp(MVN::probability(prop_i, last_j)) for every proposal i (abcsmc.cc:264-268, mvn.cc:185-196)."""
    var = np.ascontiguousarray(var, dtype=np.int32)
    inv_M = np.ascontiguousarray(inv_M, dtype=np.float64)
    pp = np.ascontiguousarray(param_prop, dtype=np.float64)
    pl = np.ascontiguousarray(param_last, dtype=np.float64)
    w = np.ascontiguousarray(w_last, dtype=np.float64)
    out = np.zeros(pp.shape[0])
    lib().oracle_mvn_kernel_sums(var.size, pp.shape[1], var.ctypes.data, inv_M.ctypes.data, float(size), pp.ctypes.data, pp.shape[0],
                                 pl.ctypes.data, w.ctypes.data, pl.shape[0], out.ctypes.data)
    return out


class OracleState:
    """Mirror of the reference's State (src/state.hh:11-133) on flat arrays."""

    def __init__(self, model, stochastic=None, keyed=False, imap_from_pop=False):
        self.model = model
        self.L = lib()
        self.h = self.L.oracle_state_create(ctypes.addressof(model.c))
        st = model.stochastic if stochastic is None else int(stochastic)
        self.L.oracle_set_mode(self.h, st, int(keyed), int(imap_from_pop))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_state_destroy(self.h)
            self.h = None

    def set_mode(self, stochastic, keyed, imap_from_pop):
        self.L.oracle_set_mode(self.h, int(stochastic), int(keyed), int(imap_from_pop))

    def _view(self, which, shape):
        p = self.L.oracle_ptr(self.h, which)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape)

    def set_param(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        assert th.size == self.model.nparam
        self.L.oracle_set_param(self.h, th.ctypes.data)

    def simulate(self, ti=0, tf=None, seed=0, stream_id=0):
        tf = self.model.T if tf is None else tf
        self.L.oracle_simulate(self.h, ti, tf, seed, stream_id)

    def init_from_particle(self, transnum):
        tn = np.ascontiguousarray(transnum, dtype=np.float64)
        self.L.oracle_init_from_particle(self.h, tn.ctypes.data)

    def recompute_transmean(self):
        self.L.oracle_recompute_transmean(self.h)

    def copy_from(self, other):
        self.L.oracle_copy_state(self.h, other.h)

    def mbp(self, init, mask=None, seed=0, stream_id=0):
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mp = mask.ctypes.data
        self.L.oracle_mbp(init.h, self.h, mp, seed, stream_id)

    def obs_values(self):
        out = np.zeros(self.model.nobs)
        self.L.oracle_obs_values(self.h, out.ctypes.data)
        return out

    def obs_loglik(self, section=-1):
        if section < 0:
            return self.L.oracle_obs_loglik(self.h)
        return self.L.oracle_obs_loglik_section(self.h, section)

    def EF_datatable(self):
        out = np.zeros(self.model.ntable)
        self.L.oracle_EF_datatable(self.h, out.ctypes.data)
        return out

    def latent_loglik(self):
        return self.L.oracle_latent_loglik(self.h)

    def graph_state(self):
        """ObservationModel::get_graph_state (obsmodel.cc:113-170), the graphs' points concatenated"""
        m = self.model
        n = sum(self.L.oracle_graph_npoints(ctypes.addressof(m.c), g) for g in range(m.graph_type.size))
        out = np.zeros(n)
        self.L.oracle_graph_state(self.h, out.ctypes.data)
        return out

    def likelihood_approx(self, theta, invT=1.0, want_state=False):
        """State::likelihood_approx (state_mvn_approx.cc:11-91); with want_state also the mean populations [T+1][A][C][D]
        and the variances carried at the start of every step [T][A][C][D]."""
        m = self.model
        th = np.ascontiguousarray(theta, dtype=np.float64)
        n = m.A * m.C * m.D
        pop = np.zeros((m.T + 1, m.A, m.C, m.D)) if want_state else None
        var = np.zeros((m.T, m.A, m.C, m.D)) if want_state else None
        L = self.L.oracle_likelihood_approx(self.h, th.ctypes.data, float(invT), pop.ctypes.data if want_state else None, var.ctypes.data if want_state else None)
        return (L, pop, var) if want_state else L

    @property
    def nsection(self):
        return self.L.oracle_nsection(self.h)

    def section_range(self, sec):
        a, b = ctypes.c_int(), ctypes.c_int()
        self.L.oracle_section_range(self.h, sec, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    @property
    def ndraw(self):
        return self.L.oracle_ndraw(self.h)

    # array views in reference order
    @property
    def transrate(self): return self._view(TRANSRATE, (self.model.J, self.model.D))
    @property
    def disc_spline(self): return self._view(DISC_SPLINE, (self.model.nspline, self.model.T))
    @property
    def susceptibility(self): return self._view(SUS, (self.model.D,))
    @property
    def areafactor(self): return self._view(AREAFACTOR, (self.model.period, self.model.A))
    @property
    def beta(self): return self._view(BETA, (self.model.S, self.model.A, self.model.T))
    @property
    def Ntime(self): return self._view(NTIME, (self.model.T, self.model.G, self.model.G))
    @property
    def pop(self): return self._view(POP, self.model.pop_shape)
    @property
    def transnum(self): return self._view(TRANSNUM, self.model.traj_shape)
    @property
    def transmean(self): return self._view(TRANSMEAN, self.model.traj_shape)
    @property
    def Imap(self): return self._view(IMAP, (self.model.T, self.model.S, self.model.A * self.model.G))
    @property
    def Idiag(self): return self._view(IDIAG, (self.model.T, self.model.S, self.model.A * self.model.G))


def bootstrap(L, invT, seed, section):
    Ls = np.ascontiguousarray(L, dtype=np.float64)
    inc = ctypes.c_double()
    back = np.zeros(Ls.size, dtype=np.uint32)
    lib().oracle_bootstrap(Ls.ctypes.data, Ls.size, invT, seed, section, ctypes.byref(inc), back.ctypes.data)
    return inc.value, back
