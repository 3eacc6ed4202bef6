"""ctypes binding of the C-ABI engine (``include/beep_b200.h``).

This is plumbing: every call goes straight to ``beep_b200/lib/libbeep_b200.so``
(hand-written sm_100a kernels).  There is no CPU fallback -- if the library is
missing or no CUDA device is present the constructors raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BEEP_B200_LIB") or os.path.join(_HERE, "lib", "libbeep_b200.so")   # override: kernel experiments only

KEEP_TRANSNUM, KEEP_TRANSMEAN, KEEP_POP = 1, 2, 4
DERIVED_TRANSRATE, DERIVED_DISC_SPLINE, DERIVED_SUSCEPTIBILITY, DERIVED_AREAFACTOR, DERIVED_BETA, DERIVED_AGEFACTOR = range(6)
TRAJ_TRANSNUM, TRAJ_TRANSMEAN, TRAJ_POP = range(3)
PERMUTE_FULL, PERMUTE_SECTION_STATE = 0, 1
INF_UPDATE, INF_DIF_UPDATE = 0, 1
KEEP = 0xFFFFFFFF

_c_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_SIGNATURES = {
    "beep_last_error": (ctypes.c_char_p, []),
    "beep_version": (ctypes.c_int, []),
    "beep_engine_create": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(_vp)]),
    "beep_engine_destroy": (None, [_vp]),
    "beep_batch_create": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint, ctypes.POINTER(_vp)]),
    "beep_batch_destroy": (None, [_vp]),
    "beep_batch_size": (ctypes.c_int, [_vp]),
    "beep_batch_bytes": (ctypes.c_int64, [_vp]),
    "beep_set_params": (ctypes.c_int, [_vp, _vp]),
    "beep_set_params_dev": (ctypes.c_int, [_vp, _vp]),
    "beep_simulate": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, _vp]),
    "beep_mbp": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_uint64, _vp]),
    "beep_replay": (ctypes.c_int, [_vp]),
    "beep_obs_loglik": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "beep_obs_loglik_dev": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "beep_obs_values": (ctypes.c_int, [_vp, _vp]),
    "beep_num_sections": (ctypes.c_int, [_vp]),
    "beep_section_range": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "beep_latent_loglik": (ctypes.c_int, [_vp, _vp]),
    "beep_replay_range": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "beep_likelihood_approx": (ctypes.c_int, [_vp, ctypes.c_double, _vp]),
    "beep_engine_has_approx": (ctypes.c_int, [_vp]),
    "beep_desc_size": (ctypes.c_int, []),
    "beep_mbp_al": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp, ctypes.c_int, _vp]),
    "beep_graph_layout": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "beep_graph_state": (ctypes.c_int, [_vp, _vp]),
    "beep_graph_statistics": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    "beep_permute": (ctypes.c_int, [_vp, _vp, ctypes.c_int]),
    "beep_accept": (ctypes.c_int, [_vp, _vp, _vp]),
    "beep_export_particle": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "beep_import_particle": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "beep_get_derived": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int64]),
    "beep_get_trajectory": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int64]),
    "beep_get_pop": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int64]),
    "beep_get_pop_all": (ctypes.c_int, [_vp, _vp]),
    "beep_set_pop_all": (ctypes.c_int, [_vp, _vp, ctypes.c_int]),
    "beep_sync": (ctypes.c_int, [_vp]),
    "beep_batch_stream": (_vp, [_vp]),
    "beep_launch_count": (ctypes.c_int64, [_vp]),
    "beep_engine_sync_scope": (ctypes.c_int, [_vp]),
    "beep_engine_chunk_tiles": (ctypes.c_int, [_vp]),
    "beep_engine_dense_mixing": (ctypes.c_int, [_vp]),
    "beep_engine_dims": (ctypes.c_int, [_vp] + [ctypes.POINTER(ctypes.c_int)] * 8),
    "beep_last_simulate_ms": (ctypes.c_double, [_vp]),
    "beep_save_sample": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, _vp, ctypes.c_double, ctypes.c_double]),
    "beep_load_sample": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, _vp, _c_dp]),
    "beep_ef_cutoff": (ctypes.c_int, [_vp, ctypes.c_int, _c_dp, _c_dp, _c_dp, _vp]),
    "beep_pas_bootstrap": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_int,
                                          _c_dp, _c_dp, _c_dp, _vp]),
    "beep_mc3_swap": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, _vp, _vp, _vp]),
    "beep_test_uniforms": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, _vp]),
    "beep_mvn_kernel_sums": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_double, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp]),
    "beep_bootstrap": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_int, _c_dp, _vp]),
    "beep_comm_unique_id": (ctypes.c_int, [_vp]),
    "beep_comm_create": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    "beep_comm_destroy": (None, [_vp]),
    "beep_comm_rank": (ctypes.c_int, [_vp]),
    "beep_comm_size": (ctypes.c_int, [_vp]),
    "beep_comm_bytes_moved": (ctypes.c_int64, [_vp, ctypes.POINTER(ctypes.c_int64)]),
    "beep_allgather_scalar": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
    "beep_allgather_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp]),
    "beep_copy_particles": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "beep_end_sec_swap": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "beep_particle_sample": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "beep_test_move_plan": (ctypes.c_int, [ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.POINTER(ctypes.c_int), _vp, ctypes.POINTER(ctypes.c_int)]),
    "beep_test_poisson": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, _vp, ctypes.c_int, _vp]),
    "beep_test_binomial": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, _vp, _vp, ctypes.c_int, _vp]),
}

_LIB = None


class BeepError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"beep_b200 error {code}: {msg}")
        self.code = code


def _preload_nccl():
    """The engine links libnccl.so.2.  In a process that also imports torch, the NCCL that torch bundles (newer
    than the system one) must be the copy that gets loaded, whichever of the two libraries comes first."""
    try:
        import glob
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            for f in glob.glob(os.path.join(d, "lib", "libnccl.so.2")):
                ctypes.CDLL(f, mode=ctypes.RTLD_GLOBAL)
                return
    except Exception:
        pass


def lib():
    """Load the engine library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        _preload_nccl()
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C beep_b200/csrc).  The BEEP engine has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise BeepError(rc, lib().beep_last_error().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """One model on one GPU (read-only tables uploaded once)."""

    def __init__(self, model, device=0):
        self.model = model
        self.L = lib()
        if self.L.beep_desc_size() != ctypes.sizeof(model.c):
            raise BeepError(-1, "BeepModelDesc of include/beep_model_fields.def and of libbeep_b200.so differ: rebuild the library")
        h = _vp()
        _check(self.L.beep_engine_create(ctypes.addressof(model.c), device, ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.beep_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def batch(self, nparticle, flags=0):
        return Batch(self, nparticle, flags)

    @property
    def nsection(self):
        return self.L.beep_num_sections(self.h)

    def section_range(self, s):
        a, b = ctypes.c_int(), ctypes.c_int()
        _check(self.L.beep_section_range(self.h, s, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def launch_count(self):
        return self.L.beep_launch_count(self.h)

    @property
    def has_approx(self):
        return bool(self.L.beep_engine_has_approx(self.h))

    @property
    def dense_mixing(self):
        return bool(self.L.beep_engine_dense_mixing(self.h))

    @property
    def chunk_tiles(self):
        return self.L.beep_engine_chunk_tiles(self.h)

    @property
    def sync_scope(self):
        return {0: "NONE", 1: "CTA", 2: "GRID"}[self.L.beep_engine_sync_scope(self.h)]


class Batch:
    """P particles resident in HBM: the batched counterpart of the reference's
    ``State`` (src/state.hh:11-133)."""

    def __init__(self, engine, nparticle, flags=0):
        self.engine = engine
        self.model = engine.model
        self.L = engine.L
        self.P = int(nparticle)
        h = _vp()
        _check(self.L.beep_batch_create(engine.h, self.P, flags, ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.engine, "h", None):
            self.L.beep_batch_destroy(self.h)
        self.h = None

    def __del__(self):
        self.close()

    @property
    def nbytes(self):
        return self.L.beep_batch_bytes(self.h)

    # -- K0 ---------------------------------------------------------------
    def set_params(self, theta):
        th = _f64(theta)
        if th.ndim == 1:
            th = np.ascontiguousarray(np.broadcast_to(th, (self.P, th.size)))
        assert th.shape == (self.P, self.model.nparam), th.shape
        _check(self.L.beep_set_params(self.h, th.ctypes.data))

    def set_params_dev(self, ptr):
        _check(self.L.beep_set_params_dev(self.h, ptr))

    # -- K1/K2 ------------------------------------------------------------
    def simulate(self, ti=0, tf=None, seed=0, stream_ids=None):
        tf = self.model.T if tf is None else tf
        sp = None
        if stream_ids is not None:
            ids = np.ascontiguousarray(stream_ids, dtype=np.uint64)
            assert ids.size == self.P
            sp = ids.ctypes.data
        _check(self.L.beep_simulate(self.h, ti, tf, seed, sp))

    # -- K3 ---------------------------------------------------------------
    def mbp(self, init, mask=None, seed=0, stream_ids=None, inf_update=INF_DIF_UPDATE):
        mp = sp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if mask.ndim == 2:
                mask = np.ascontiguousarray(np.broadcast_to(mask, (self.P,) + mask.shape))
            assert mask.shape == (self.P, self.model.T, self.model.A)
            mp = mask.ctypes.data
        if stream_ids is not None:
            ids = np.ascontiguousarray(stream_ids, dtype=np.uint64)
            sp = ids.ctypes.data
        _check(self.L.beep_mbp(init.h, self.h, mp, inf_update, seed, sp))

    def replay(self):
        _check(self.L.beep_replay(self.h))

    # -- K4 / K5 ----------------------------------------------------------
    def replay_range(self, ti, tf):
        _check(self.L.beep_replay_range(self.h, ti, tf))

    def obs_loglik(self, section=-1, with_tables=False):
        Lh = np.zeros(self.P)
        if with_tables:
            tab = np.zeros((self.P, self.model.ntable))
            _check(self.L.beep_obs_loglik(self.h, section, Lh.ctypes.data, tab.ctypes.data))
            return Lh, tab
        _check(self.L.beep_obs_loglik(self.h, section, Lh.ctypes.data, None))
        return Lh

    def obs_values(self):
        out = np.zeros((self.P, self.model.nobs))
        _check(self.L.beep_obs_values(self.h, out.ctypes.data))
        return out

    def latent_loglik(self):
        out = np.zeros(self.P)
        _check(self.L.beep_latent_loglik(self.h, out.ctypes.data))
        return out

    # -- K6 ---------------------------------------------------------------
    def graph_state(self):
        """ObservationModel::get_graph_state (obsmodel.cc:113-170) of every particle: [P][npoint]"""
        ng, np_ = ctypes.c_int(), ctypes.c_int()
        _check(self.L.beep_graph_layout(self.engine.h, ctypes.byref(ng), ctypes.byref(np_), None))
        out = np.zeros((self.P, np_.value))
        _check(self.L.beep_graph_state(self.h, out.ctypes.data))
        return out

    def likelihood_approx(self, invT=1.0):
        """State::likelihood_approx (state_mvn_approx.cc:11-91) of every particle, parameters from set_params."""
        out = np.zeros(self.P)
        _check(self.L.beep_likelihood_approx(self.h, float(invT), out.ctypes.data))
        return out

    def permute(self, src_of, what=PERMUTE_FULL):
        m = np.ascontiguousarray(src_of, dtype=np.uint32)
        assert m.size == self.P
        _check(self.L.beep_permute(self.h, m.ctypes.data, what))

    def accept_from(self, prop, accept):
        a = np.ascontiguousarray(accept, dtype=np.uint8)
        assert a.size == self.P
        _check(self.L.beep_accept(self.h, prop.h, a.ctypes.data))

    # -- particles ----------------------------------------------------------
    def export_particle(self, slot):
        m = self.model
        th = np.zeros(m.nparam)
        tn = np.zeros(m.traj_shape)
        _check(self.L.beep_export_particle(self.h, slot, th.ctypes.data, tn.ctypes.data))
        return th, tn

    def import_particle(self, slot, theta, transnum):
        th, tn = _f64(theta), _f64(transnum)
        _check(self.L.beep_import_particle(self.h, slot, th.ctypes.data, tn.ctypes.data))

    def save_sample(self, slot, prefix, param_names, EF, start_time=0.0):
        """State::save (state.cc:709-739): <prefix>_parameter.csv and <prefix>_transition.csv"""
        arr = (ctypes.c_char_p * len(param_names))(*[n.encode() for n in param_names])
        _check(self.L.beep_save_sample(self.h, slot, prefix.encode(), ctypes.cast(arr, ctypes.c_void_p), float(EF), float(start_time)))

    def load_sample(self, slot, prefix, param_names=None):
        """State::load (state.cc:743-790); returns the stored error function.  Call replay() afterwards."""
        ef = ctypes.c_double()
        arr = None
        if param_names is not None:
            arr = ctypes.cast((ctypes.c_char_p * len(param_names))(*[n.encode() for n in param_names]), ctypes.c_void_p)
        _check(self.L.beep_load_sample(self.h, slot, prefix.encode(), arr, ctypes.byref(ef)))
        return ef.value

    # -- read-back ----------------------------------------------------------
    def derived(self, slot, which):
        m = self.model
        shape = {DERIVED_TRANSRATE: (m.J, m.D), DERIVED_DISC_SPLINE: (m.nspline, m.T), DERIVED_SUSCEPTIBILITY: (m.D,),
                 DERIVED_AREAFACTOR: (m.period, m.A), DERIVED_BETA: (m.S, m.A, m.T), DERIVED_AGEFACTOR: (m.T, m.G)}[which]
        out = np.zeros(shape)
        _check(self.L.beep_get_derived(self.h, slot, which, out.ctypes.data, out.size))
        return out

    def trajectory(self, slot, which=TRAJ_TRANSNUM):
        m = self.model
        out = np.zeros(m.pop_shape if which == TRAJ_POP else m.traj_shape)
        _check(self.L.beep_get_trajectory(self.h, slot, which, out.ctypes.data, out.size))
        return out

    def pop(self, slot):
        m = self.model
        out = np.zeros((m.A, m.C, m.D))
        _check(self.L.beep_get_pop(self.h, slot, out.ctypes.data, out.size))
        return out

    def pop_all(self):
        m = self.model
        out = np.zeros((self.P, m.A, m.C, m.D))
        _check(self.L.beep_get_pop_all(self.h, out.ctypes.data))
        return out

    def set_pop_all(self, pop, cur_step):
        p = _f64(pop)
        assert p.size == self.P * self.model.A * self.model.C * self.model.D
        _check(self.L.beep_set_pop_all(self.h, p.ctypes.data, int(cur_step)))

    def sync(self):
        _check(self.L.beep_sync(self.h))

    @property
    def stream(self):
        return self.L.beep_batch_stream(self.h)

    @property
    def last_simulate_ms(self):
        return self.L.beep_last_simulate_ms(self.h)


def comm_unique_id():
    """128 bytes of an ncclUniqueId (create on one rank, hand to the others)."""
    buf = ctypes.create_string_buffer(128)
    _check(lib().beep_comm_unique_id(buf))
    return buf.raw


class Comm:
    """The collectives of the particle methods inside the C-ABI library (NCCL, device to device): the
    replacement of the reference's Mpi object for the hot path (src/mpi.cc:81-165,215-389)."""

    def __init__(self, engine, unique_id, nranks, rank):
        self.engine = engine
        self.L = engine.L
        self.nranks, self.rank = int(nranks), int(rank)
        h = _vp()
        idbuf = ctypes.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(self.L.beep_comm_create(engine.h, idbuf, self.nranks, self.rank, ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.beep_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def allgather(self, local):
        """per-particle scalars (or rows) of every rank, in rank order"""
        loc = _f64(local)
        out = np.zeros((self.nranks,) + loc.shape).reshape((self.nranks * loc.shape[0],) + loc.shape[1:])
        _check(self.L.beep_allgather_scalar(self.h, loc.ctypes.data, loc.size, out.ctypes.data))
        return out

    def copy_particles(self, batch, partcopy_all):
        pc = np.ascontiguousarray(partcopy_all, dtype=np.uint32)
        _check(self.L.beep_copy_particles(self.h, batch.h, pc.ctypes.data, pc.size))

    def end_sec_swap(self, batch, backpart_all):
        bp = np.ascontiguousarray(backpart_all, dtype=np.uint32)
        _check(self.L.beep_end_sec_swap(self.h, batch.h, bp.ctypes.data, bp.size))

    def particle_sample(self, batch, backpart):
        """backpart: [nsec][Ntot].  Returns (theta, transnum) on rank 0, None elsewhere."""
        bp = np.ascontiguousarray(backpart, dtype=np.uint32)
        m = batch.model
        th, tn = np.zeros(m.nparam), np.zeros(m.traj_shape)
        _check(self.L.beep_particle_sample(self.h, batch.h, bp.ctypes.data, bp.shape[0], bp.shape[1], th.ctypes.data, tn.ctypes.data))
        return (th, tn) if self.rank == 0 else None

    @property
    def bytes_moved(self):
        r = ctypes.c_int64()
        s = self.L.beep_comm_bytes_moved(self.h, ctypes.byref(r))
        return s, r.value


def particle_sample(batch, backpart):
    """Mpi::particle_sample on one GPU (no communicator)."""
    bp = np.ascontiguousarray(backpart, dtype=np.uint32)
    m = batch.model
    th, tn = np.zeros(m.nparam), np.zeros(m.traj_shape)
    _check(lib().beep_particle_sample(None, batch.h, bp.ctypes.data, bp.shape[0], bp.shape[1], th.ctypes.data, tn.ctypes.data))
    return th, tn


def move_plan(kind, vec_all, N, rank):
    """test hook: (local[N], out [(peer, slot)], in [(peer, [slots])]) of one rank; kind 0 copy_particles, 1 end_sec_swap"""
    v = np.ascontiguousarray(vec_all, dtype=np.uint32)
    local = np.zeros(N, dtype=np.uint32)
    outp = np.zeros(2 * v.size + 2, dtype=np.int32)
    inf = np.zeros(3 * v.size + 3, dtype=np.int32)
    nout, nin = ctypes.c_int(), ctypes.c_int()
    _check(lib().beep_test_move_plan(kind, v.ctypes.data, v.size, N, rank, local.ctypes.data, outp.ctypes.data, ctypes.byref(nout), inf.ctypes.data, ctypes.byref(nin)))
    out = [(int(outp[2 * k]), int(outp[2 * k + 1])) for k in range(nout.value)]
    ins, w = [], 0
    while w < nin.value:
        peer, n = int(inf[w]), int(inf[w + 1])
        ins.append((peer, [int(x) for x in inf[w + 2:w + 2 + n]]))
        w += 2 + n
    return local, out, ins


def mbp_al(mode, invT, EFcut, EF_i, EF_p, Pr_i, Pr_p):
    """Mbp::get_al (mbp.cc:334-361); mode 0 = CUTOFF, 1 = INVT"""
    a = [np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64) for x in (EF_i, EF_p, Pr_i, Pr_p)]
    out = np.zeros(a[0].size)
    _check(lib().beep_mbp_al(int(mode), float(invT), float(EFcut), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, a[0].size, out.ctypes.data))
    return out


def graph_statistics(gs):
    """Output::get_statistic (output.cc:3089-3123) per graph point over the samples: mean, CI min, CI max"""
    gs = np.ascontiguousarray(gs, dtype=np.float64)
    P, n = gs.shape
    mean, lo, hi = np.zeros(n), np.zeros(n), np.zeros(n)
    _check(lib().beep_graph_statistics(gs.ctypes.data, P, n, mean.ctypes.data, lo.ctypes.data, hi.ctypes.data))
    return mean, lo, hi


def mvn_kernel_sums(engine, var, inv_M, size, param_prop, param_last, w_last):
    """ABCSMC::calculate_particle_weight's sum over the previous generation (abcsmc.cc:264-268) for every
    proposal: SUM_j w_last[j] exp(MVN::probability(prop_i, last_j)) -- beep_mvn_kernel_sums."""
    var = np.ascontiguousarray(var, dtype=np.int32)
    inv_M = _f64(inv_M)
    pp, pl, w = _f64(param_prop), _f64(param_last), _f64(w_last)
    if pp.ndim != 2 or pl.ndim != 2 or pp.shape[1] != pl.shape[1] or w.size != pl.shape[0] or inv_M.size != var.size ** 2:
        raise ValueError("inconsistent shapes")
    out = np.zeros(pp.shape[0])
    _check(engine.L.beep_mvn_kernel_sums(engine.h, var.size, pp.shape[1], var.ctypes.data, inv_M.ctypes.data, float(size),
                                         pp.ctypes.data, pp.shape[0], pl.ctypes.data, w.ctypes.data, pl.shape[0], out.ctypes.data))
    return out


def bootstrap(L_all, invT, seed, section):
    """PMCMC::bootstrap (src/pmcmc.cc:211-268) on the gathered weights."""
    Ls = _f64(L_all)
    inc = ctypes.c_double()
    back = np.zeros(Ls.size, dtype=np.uint32)
    _check(lib().beep_bootstrap(Ls.ctypes.data, Ls.size, invT, seed, section, ctypes.byref(inc), back.ctypes.data))
    return inc.value, back


def test_poisson(seed, stream_id, lam):
    lam = _f64(lam)
    out = np.zeros(lam.size, dtype=np.int32)
    _check(lib().beep_test_poisson(seed, stream_id, lam.ctypes.data, lam.size, out.ctypes.data))
    return out


def test_binomial(seed, stream_id, p, n):
    p = _f64(p)
    n = np.ascontiguousarray(n, dtype=np.int32)
    out = np.zeros(p.size, dtype=np.int32)
    _check(lib().beep_test_binomial(seed, stream_id, p.ctypes.data, n.ctypes.data, p.size, out.ctypes.data))
    return out
