/* beep_b200.h -- C ABI of the B200-native BEEP latent-state engine.
 *
 * The reference (ScottishCovidResponse/BEEP @ d70908f) has no plugin or FFI
 * layer; its seam is the public surface of three C++ classes that every
 * inference driver calls directly:
 *     State              src/state.hh:11-133
 *     Mbp                src/mbp.hh:15-82
 *     ObservationModel   src/obsmodel.hh:8-39
 * plus the three particle movers of Mpi (src/mpi.hh:16-24).  Each entry point
 * below names the reference member function(s) whose body it replaces; the
 * drivers (mbp.cc, pas.cc, abcsmc.cc, abcmbp.cc, pmcmc.cc, mc3.cc, ml.cc)
 * keep their host loops and call these with a BATCH of particles instead of
 * one State at a time (see INTEGRATION.md for the reference-side adapter).
 *
 * Conventions
 *   - plain C types only; no torch/STL types cross this boundary;
 *   - every function returns BEEP_OK (0) or a negative BeepStatus and never
 *     calls exit() (the reference's emsg() exits, utils.cc:296-316);
 *     beep_last_error() gives the message for the calling thread;
 *   - the engine owns all device memory; the caller owns host arrays;
 *   - pointers named *_host are host memory (any, pinned is faster); pointers
 *     named *_dev are device memory on the engine's GPU;
 *   - particle trajectories cross the boundary in exactly Particle::transnum
 *     order [sett][area][trans][dp] (struct.hh:176-188, state.cc:728-737);
 *   - there is no global RNG: every draw is a pure function of
 *     (seed, stream_id[p], sett, area, dp, draw index) -- see beep_rng.h;
 *   - calls on one batch are asynchronous on the batch's CUDA stream and are
 *     ordered; any function returning data to the host synchronises.
 *   - there is no CPU fallback: without a CUDA device engine creation fails.
 */
#ifndef BEEP_B200_H
#define BEEP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ */
/* Model description (flattened Details + Data + Model)                      */
/* ------------------------------------------------------------------------ */
typedef struct BeepModelDesc {
#define BEEP_I32(name)  int32_t name;
#define BEEP_F64(name)  double name;
#define BEEP_AI32(name) const int32_t *name; int64_t n_##name;
#define BEEP_AF64(name) const double *name; int64_t n_##name;
#include "beep_model_fields.def"
#undef BEEP_I32
#undef BEEP_F64
#undef BEEP_AI32
#undef BEEP_AF64
} BeepModelDesc;

/* Sentinels compared as doubles, reference src/consts.hh:159-161 */
#define BEEP_UNSET   999999999.0
#define BEEP_THRESH  999999998.0
#define BEEP_UNKNOWN 999999997.0

/* ObsModelFunc, reference src/consts.hh:40 */
enum { BEEP_OBS_NORMAL = 0, BEEP_OBS_NORMAL_PERCENT = 1, BEEP_OBS_POISSON = 2, BEEP_OBS_NEGBIN = 3 };

typedef enum BeepStatus {
	BEEP_OK = 0,
	BEEP_ERR_INVALID = -1,      /* bad argument / inconsistent model description   */
	BEEP_ERR_UNSUPPORTED = -2,  /* feature outside the hot-path scope              */
	BEEP_ERR_CUDA = -3,         /* CUDA runtime error (message has the detail)     */
	BEEP_ERR_NO_DEVICE = -4,    /* no CUDA device: there is no CPU fallback        */
	BEEP_ERR_STATE = -5,        /* call order (e.g. simulate before set_params)    */
	BEEP_ERR_NUMERIC = -6       /* NaN / non-convergence (reference emsgEC sites)  */
} BeepStatus;

typedef struct BeepEngine BeepEngine;
typedef struct BeepBatch BeepBatch;

/* Batch allocation flags */
enum {
	BEEP_KEEP_TRANSNUM  = 1,  /* store transnum[P][T][A][J][D] (MBP chains, samples)  */
	BEEP_KEEP_TRANSMEAN = 2,  /* store transmean likewise (MBP initial state, K5)     */
	BEEP_KEEP_POP       = 4   /* store pop[P][T][A][C][D] (outputs / parity tests)    */
};

/* Selectors for beep_get_derived (outputs of State::set_param, state.hh:56-64) */
enum {
	BEEP_DERIVED_TRANSRATE = 0,      /* [J][D]                                        */
	BEEP_DERIVED_DISC_SPLINE = 1,    /* [nspline][T]                                  */
	BEEP_DERIVED_SUSCEPTIBILITY = 2, /* [D]                                           */
	BEEP_DERIVED_AREAFACTOR = 3,     /* [period][A]                                   */
	BEEP_DERIVED_BETA = 4,           /* [S][A][T]                                     */
	BEEP_DERIVED_AGEFACTOR = 5       /* [T][G] g with Ntime[t][a][b]=N0[a][b] g_a g_b  */
};

/* Selectors for beep_get_trajectory */
enum { BEEP_TRAJ_TRANSNUM = 0, BEEP_TRAJ_TRANSMEAN = 1, BEEP_TRAJ_POP = 2 };

/* InfUpdate, reference src/consts.hh:131 */
enum { BEEP_INF_UPDATE = 0, BEEP_INF_DIF_UPDATE = 1 };

/* beep_permute `what` */
enum { BEEP_PERMUTE_FULL = 0, BEEP_PERMUTE_SECTION_STATE = 1 };

const char *beep_last_error(void);
int beep_version(void);
/* sizeof(BeepModelDesc) the library was built with: a caller compiled against another beep_model_fields.def must not
 * pass its description (the C++ wrapper checks)                                                        */
int beep_desc_size(void);

/* Replaces: construction of Data/Model tables per MPI rank (mpi.cc:26-77
 * broadcast) -- uploads the read-only tables once to `device`.              */
int beep_engine_create(const BeepModelDesc *desc, int device, BeepEngine **out);
void beep_engine_destroy(BeepEngine *eng);

/* Replaces: `vector<State> particle` / `vector<Particle> part` owned by the
 * drivers (pmcmc.cc:60-66, abcmbp.cc:41-48, mc3.cc:99-118).                 */
int beep_batch_create(BeepEngine *eng, int nparticle, unsigned flags, BeepBatch **out);
void beep_batch_destroy(BeepBatch *batch);
int beep_batch_size(const BeepBatch *batch);
/* bytes of device memory held by the batch (memory-frontier planning)       */
int64_t beep_batch_bytes(const BeepBatch *batch);

/* K0.  Replaces State::set_param (state.cc:66-91) for all P particles.
 * theta is [P][nparam], raw (pre-Dirichlet) parameter values.               */
int beep_set_params(BeepBatch *batch, const double *theta_host);
int beep_set_params_dev(BeepBatch *batch, const double *theta_dev);

/* K1+K2 (+K4 accumulation).  Replaces State::simulate(ti,tf)
 * (state.cc:596-638): tau-leap from step ti to tf for every particle,
 * accumulating the observation sums of those steps.  ti == 0 resets the
 * population to pop_init (state.cc:600) and clears the accumulators.
 * stream_id_host: [P] per-particle stream ids or NULL (= slot index).       */
int beep_simulate(BeepBatch *batch, int ti, int tf, uint64_t seed, const uint64_t *stream_id_host);

/* K3.  Replaces Mbp::mbp (mbp.cc:252-330): proposes a trajectory for every
 * particle of `prop` (whose parameters were set with beep_set_params) coupled
 * to the same slot of `init`, which must hold transnum and transmean.
 * simu_mask_host: [P][T][A] bytes, 1 = SIMU, 0 = MBP (consts.hh:99), or NULL
 * for pure MBP.  inf_update is accepted for signature parity; the engine
 * always derives the infectivity map from the proposed population.          */
int beep_mbp(BeepBatch *init, BeepBatch *prop, const uint8_t *simu_mask_host, int inf_update,
             uint64_t seed, const uint64_t *stream_id_host);

/* Replaces State::initialise_from_particle (state.cc:294-322) and the
 * set_transmean sweep of Mbp::parameter_proposal (mbp.cc:470-476): with the
 * batch's current parameters and stored transnum, recompute pop, transmean
 * and the observation sums.                                                 */
int beep_replay(BeepBatch *batch);
/* the steps [ti, tf) only: the batch then stands at step tf and beep_simulate(batch, tf, T, ...) continues from the replayed
 * population -- State::load + State::simulate(pred_start, ndivision) of mode=prediction (simulate.cc:141-165)          */
int beep_replay_range(BeepBatch *batch, int ti, int tf);

/* K4.  Replaces ObservationModel::calculate (obsmodel.cc:22-35, section < 0),
 * calculate_section (:39-58, section >= 0) and get_EF_datatable (:309-322).
 * L_host[P] receives the log observation probability (EF = -2 L,
 * state.cc:95-98); EF_table_host[P][ntable] may be NULL.                    */
int beep_obs_loglik(BeepBatch *batch, int section, double *L_host, double *EF_table_host);
int beep_obs_loglik_dev(BeepBatch *batch, int section, double *L_dev, double *EF_table_dev);
/* Replaces ObservationModel::get_obs_value (obsmodel.cc:62-70): [P][nobs]   */
int beep_obs_values(BeepBatch *batch, double *obs_value_host);
/* number of PMCMC sections and their step ranges (obsmodel.cc:326-348)      */
int beep_num_sections(const BeepEngine *eng);
int beep_section_range(const BeepEngine *eng, int section, int *ti, int *tf);

/* K5.  Replaces State::likelihood (state.cc:812-826): Li_host[P]            */
int beep_latent_loglik(BeepBatch *batch, double *Li_host);

/* K6.  Replaces Mpi::copy_particles (mpi.cc:81-165, what = FULL) and
 * Mpi::end_sec_swap (mpi.cc:215-327, what = SECTION_STATE) within one GPU:
 * slot p takes the state of slot src_of_host[p] (0xFFFFFFFF = keep).        */
int beep_permute(BeepBatch *batch, const uint32_t *src_of_host, int what);
/* Swap the roles of two batches for the slots with accept_host[p] != 0:
 * Mbp::swap_initial_propose_state (mbp.cc:243-248) per particle.            */
int beep_accept(BeepBatch *init, BeepBatch *prop, const uint8_t *accept_host);

/* Replaces State::create_particle (state.cc:326-335) / State::save
 * (:709-739): theta[nparam], transnum[T][A][J][D] as doubles.               */
int beep_export_particle(BeepBatch *batch, int slot, double *theta_host, double *transnum_host);
/* Replaces State::load / initialise_from_particle (state.cc:294-322,743-808):
 * stores theta + transnum in the slot; call beep_set_params is NOT needed
 * afterwards for that slot, but beep_replay is (it is batch-wide).          */
int beep_import_particle(BeepBatch *batch, int slot, const double *theta_host, const double *transnum_host);

/* Read-back used by the parity tests and by output code on the host.        */
int beep_get_derived(BeepBatch *batch, int slot, int which, double *out_host, int64_t n);
int beep_get_trajectory(BeepBatch *batch, int slot, int which, double *out_host, int64_t n);
/* population at the batch's current step, [A][C][D]                         */
int beep_get_pop(BeepBatch *batch, int slot, double *out_host, int64_t n);

/* Section state of the whole batch, [P][A][C][D]: what Mpi::end_sec_swap ships between ranks
 * (mpi.cc:286-288; Imap/Idiag are re-derived from pop).  set also declares the step the batch
 * now stands at, so that the next beep_simulate(ti, ...) continues from it.                    */
int beep_get_pop_all(BeepBatch *batch, double *pop_host);
int beep_set_pop_all(BeepBatch *batch, const double *pop_host, int cur_step);

int beep_sync(BeepBatch *batch);
/* the batch's cudaStream_t, for callers that time or enqueue on it          */
void *beep_batch_stream(BeepBatch *batch);
/* scope of the per-step barrier the trajectory kernel uses for this model:
 * 0 = none (areas uncoupled), 1 = CTA (a CTA owns a particle), 2 = cooperative grid */
int beep_engine_sync_scope(const BeepEngine *eng);
/* the dimensions the engine was created with (any pointer may be NULL) */
int beep_engine_dims(const BeepEngine *eng, int *T, int *A, int *C, int *J, int *D, int *nparam, int *nobs, int *ntable);
/* tiles (consecutive areas of one particle-step) a warp of the grid-spanning trajectory kernel works at once:
 * 4 or 8 when the chunk kernel (integer populations, pooled draws) serves the model, 0 when the
 * one-tile-per-warp kernel does */
int beep_engine_chunk_tiles(const BeepEngine *eng);
/* 1 when the area-mixing contraction (state.cc:204-216) runs as a batched FP64 tensor-core GEMM per step (dense
 * mixing matrix: at least 64 off-diagonal entries per area, or BEEP_DENSE_M=1), 0 when it is the sparse gather */
int beep_engine_dense_mixing(const BeepEngine *eng);
/* number of engine kernel launches since creation (bench accounting)        */
int64_t beep_launch_count(const BeepEngine *eng);
/* last kernel-only duration of beep_simulate in ms (CUDA events on the
 * batch stream), -1 if none.                                                */
double beep_last_simulate_ms(const BeepBatch *batch);

/* PMCMC bootstrap step, pmcmc.cc:211-268, on the weights of ALL ranks
 * (already all-gathered by the caller): computes the log-likelihood
 * increment and the ancestor vector with the stay-in-place rule.  Pure host
 * arithmetic on O(Ntot) scalars, replicated on every rank from a shared
 * (seed, section) key so no broadcast is needed.                            */
int beep_bootstrap(const double *L_all_host, int ntot, double invT, uint64_t seed, int section,
                   double *loglik_inc, uint32_t *backpart_host);

/* Posterior sample files: State::save / State::load
 * (state.cc:709-790) straight from / into a slot of a batch.  Writes and reads
 * <prefix>_parameter.csv and <prefix>_transition.csv byte-compatibly with the
 * reference's Posterior/samples/sampleN_* files (same number formatting, same
 * [sett][area][trans][dp] order, times = details.division_time).  param_names:
 * model.param[th].name, [nparam]; start_time: details.start for
 * time_format = "number", else 0.  After beep_load_sample call beep_replay to
 * rebuild populations and means (State::initialise_from_particle).            */
int beep_save_sample(BeepBatch *batch, int slot, const char *file_prefix, const char *const *param_names, double EF, double start_time);
int beep_load_sample(BeepBatch *batch, int slot, const char *file_prefix, const char *const *param_names, double *EF_out);

/* ------------------------------------------------------------------------ */
/* Collectives of the particle methods: NCCL over NVLink, device to device   */
/* ------------------------------------------------------------------------ */
/* One BeepComm per GPU / rank replaces the reference's Mpi object (src/mpi.hh:16-60) for the hot path.  Global
 * particle i lives on rank i / N (N = particles per rank), the reference's core = i/N (mpi.cc:98-101).
 * beep_comm_unique_id writes the 128 bytes of an ncclUniqueId on ONE rank; the caller hands them to the other
 * ranks by whatever it has (MPI_Bcast in the reference's host code, a file, torch.distributed in the bench).   */
typedef struct BeepComm BeepComm;
int beep_comm_unique_id(void *id128);
int beep_comm_create(BeepEngine *eng, const void *id128, int nranks, int rank, BeepComm **out);
void beep_comm_destroy(BeepComm *comm);
int beep_comm_rank(const BeepComm *comm);
int beep_comm_size(const BeepComm *comm);
/* bytes this rank has sent (returned) and received (*recv) in particle movements so far */
int64_t beep_comm_bytes_moved(const BeepComm *comm, int64_t *recv);
/* Replaces mpi.gather + decision on rank 0 + mpi.bcast of per-particle scalars (abcmbp.cc:141, pas.cc:133,
 * pmcmc.cc:219,263, mc3.cc:170-175; ABC-SMC's (theta, EF, w) exchange mpi.cc:169-211): every rank gets every
 * rank's n_local values, in rank order.  _scalar: host arrays; _dev: device arrays on `stream` (NULL = the
 * communicator's own).                                                                                         */
int beep_allgather_scalar(BeepComm *comm, const double *local_host, int n_local, double *all_host);
int beep_allgather_dev(BeepComm *comm, const double *local_dev, int n_local, double *all_dev, void *stream);
/* Replaces Mpi::copy_particles (mpi.cc:81-165): global particle i continues as a copy of partcopy_all[i]
 * (0xFFFFFFFF = keep; the output of beep_ef_cutoff / beep_pas_bootstrap / beep_mc3_swap).  Whole particles move
 * HBM to HBM as their per-particle slabs (grouped ncclSend / ncclRecv); copies within a rank are one device
 * permutation.  The state arrives complete: no replay is needed afterwards.                                    */
int beep_copy_particles(BeepComm *comm, BeepBatch *batch, const uint32_t *partcopy_all_host, int ntot);
/* Replaces Mpi::end_sec_swap (mpi.cc:215-327): particle p continues from the section state of backpart_all[p]
 * (the output of beep_bootstrap).  Only the DISTINCT ancestors that live on another rank travel.               */
int beep_end_sec_swap(BeepComm *comm, BeepBatch *batch, const uint32_t *backpart_all_host, int ntot);
/* Replaces Mpi::particle_sample (mpi.cc:331-389): one filtered trajectory assembled backwards through the
 * ancestry backpart_host[nsec][Ntot] (row s = beep_bootstrap's output after section s).  Rank 0 receives
 * theta_host[nparam] and transnum_host[T][A][J][D]; comm may be NULL on a single GPU.                          */
int beep_particle_sample(BeepComm *comm, BeepBatch *batch, const uint32_t *backpart_host, int nsec, int ntot, double *theta_host, double *transnum_host);
/* test hook: the send / receive lists one rank derives (kind 0 = copy_particles, 1 = end_sec_swap)             */
int beep_test_move_plan(int kind, const uint32_t *vec_all, int ntot, int N, int rank, uint32_t *local, int32_t *out_pairs, int *nout,
                        int32_t *in_flat, int *nin_words);

/* Resampling decisions of the particle methods, on
 * scalars that every rank has all-gathered: pure host arithmetic, replicated
 * on every rank from a counter-based stream keyed by (seed, generation/sweep)
 * -- the reference decides on rank 0 with ran() and broadcasts.  partcopy[i] /
 * order[i] are the argument of Mpi::copy_particles (mpi.cc:81-165):
 * the particle that position i continues with, 0xFFFFFFFF = keep.
 *   beep_ef_cutoff      ABCMBP::EF_cutoff   abcmbp.cc:135-217
 *   beep_pas_bootstrap  PAS::bootstrap      pas.cc:127-209   (invT_final < 0: unset)
 *   beep_mc3_swap       MC3::swap_states    mc3.cc:153-217   (ntr/nac: counts to add per position) */
/* Mbp::get_al (mbp.cc:334-361) for n chains at once: al = min(1, exp(Pr_p - Pr_i)) if EF_p < EFcut else 0 (mode 0, CUTOFF);
 * al = min(1, exp(-invT (EF_p - EF_i)/2 + Pr_p - Pr_i)) (mode 1, INVT).  EF = -2 L from beep_obs_loglik, Pr = model.prior. */
int beep_mbp_al(int mode, double invT, double EFcut, const double *EF_i, const double *EF_p, const double *Pr_i, const double *Pr_p, int n, double *al_out);
int beep_ef_cutoff(const double *EF_all_host, int ntot, double *EFcut, double *EFmin, double *EFmax, uint32_t *partcopy_host);
int beep_pas_bootstrap(const double *EF_all_host, int nrun, int npart, double invT, double quench_factor, double invT_final,
                       uint64_t seed, int gen, double *DinvT, double *EFmin, double *EFmax, uint32_t *partcopy_host);
int beep_mc3_swap(const uint32_t *run_all_host, const double *invT_all_host, const double *EF_all_host, int ntot, uint64_t seed, int sweep,
                  uint32_t *order_host, uint32_t *ntr_inc_host, uint32_t *nac_inc_host);
/* test hook: the uniforms the decisions above consume (key = generation/sweep, tag 0x9A5 PAS / 0x3C3 MC3) */
int beep_test_uniforms(uint64_t seed, uint32_t key, uint32_t tag, int n, double *out_host);

/* ABC-SMC importance weights.  Replaces the loop of
 * ABCSMC::calculate_particle_weight (abcsmc.cc:260-271) over the previous
 * generation, with MVN::probability (mvn.cc:185-196) inside, for ALL proposals
 * of a generation in one call:
 *   sums_out[i] = SUM_j w_last[j] exp(-1/2 (d/size)' inv_M (d/size)),
 *   d[v] = param_prop[i][var[v]] - param_last[j][var[v]],  v < nvar.
 * inv_M [nvar][nvar], size and var are MVN::inv_M, MVN::size and MVN::var of the
 * reference's proposal kernel (mvn.hh:21-23,52); the caller finishes the weight
 * as exp(model.prior(param_prop_i)) / sums_out[i] (abcsmc.cc:270).  All arrays
 * are host arrays; parameters are rows of nparam doubles.  Agreement with the
 * reference's own double loop: 1e-10 relative (measured ~1e-12).             */
int beep_mvn_kernel_sums(BeepEngine *eng, int nvar, int nparam, const int *var, const double *inv_M, double size,
                         const double *param_prop, int N, const double *param_last, const double *w_last, int Nlast,
                         double *sums_out);

/* MVN-approximate likelihood of mode=map.  Replaces
 * State::likelihood_approx (state_mvn_approx.cc:11-91: the mean population and its
 * covariance carried through the steps, a Kalman-type update at every observation
 * time, apply_observation_fast :688-817) as ML::generate_samples calls it for each
 * sample of a CMA-ES generation (ml.cc:389-413) and ML's Hessian code for O(nvar^2)
 * parameter vectors (ml.cc:540-575): here for ALL particles of the batch in one
 * launch, with the parameters given by beep_set_params.  L_host[P] receives the
 * approximate log-likelihood at inverse temperature invT (-1e9 where the reference
 * returns -LARGE).  Observation slices are the reference's generate_obs_slice
 * (obsmodel.cc:357-385).  BEEP_ERR_UNSUPPORTED where the reference itself refuses
 * the model (several strains, transitions != non-susceptible compartments,
 * observations on the susceptible compartment) or the description lacks table_type.
 * Agreement with the reference's own likelihood_approx: 1e-9 relative (its AVX
 * sums run in another order; measured 1e-12 on well-conditioned parameters).   */
int beep_likelihood_approx(BeepBatch *batch, double invT, double *L_host);
int beep_engine_has_approx(const BeepEngine *eng);

/* Posterior state summaries straight from the device buffers.
 * beep_graph_state replaces ObservationModel::get_graph_state (obsmodel.cc:113-170) for
 * every particle of a batch that keeps its trajectory (and pop, when a graph plots
 * compartments): out_host [P][npoint], the points of the graphs of data.graph
 * concatenated (beep_graph_layout: point_off[ngraph+1]).  beep_graph_statistics is
 * Output::get_statistic (output.cc:3089-3123) per point over the P samples -- the Mean /
 * 95% CI min / 95% CI max columns of Posterior/state/<graph>.csv (output.cc:1109-1125).  */
int beep_graph_layout(const BeepEngine *eng, int *ngraph, int *npoint, int *point_off);
int beep_graph_state(BeepBatch *batch, double *out_host);
int beep_graph_statistics(const double *gs_host, int P, int npoint, double *mean, double *CImin, double *CImax);

/* Test hooks: the device samplers of beep_rng.h on n independent streams
 * (seed, stream_id, i), for the host/device bit-exactness tests.            */
int beep_test_poisson(uint64_t seed, uint64_t stream_id, const double *lam_host, int n, int32_t *out_host);
int beep_test_binomial(uint64_t seed, uint64_t stream_id, const double *p_host, const int32_t *n_host, int n, int32_t *out_host);

#ifdef __cplusplus
}
#endif
#endif /* BEEP_B200_H */
