/* dev_types.h -- device-side views of the model tables and of a particle batch.
 *
 * DevModel is BeepModelDesc (include/beep_model_fields.def) after the engine has
 * uploaded it and built the derived read-only indices (transposed mixing matrix,
 * observation cell index, sections).  DevBatch is the per-batch state in HBM.
 *
 * HBM layout (P = particles of the batch; "tile" = one (particle, area) pair):
 *   theta, pd            [P][nparam]           raw / Dirichlet-corrected parameters
 *   transrate            [P][J][D]             State::transrate      (state.hh:58)
 *   disc_spline          [P][nspline][T]       State::disc_spline    (state.hh:57)
 *   sus                  [P][D]                State::susceptibility (state.hh:59)
 *   areafactor           [P][period][A]        State::areafactor     (state.hh:60)
 *   betaR                [P][S][nR][T]         State::beta per R-spline group (areas of
 *                                              one group share the series, model.cc:1343)
 *   infco                [P][C]                infectivity of each compartment
 *   etab                 [P][nefoi][T]         external force of infection per individual
 *   pop_cur              [P][A][CDp]           population [C][D] at the batch's current step,
 *                                              tile stride padded to an even count
 *   dinf                 [2][P][S][A][G]       infectivity by (area, age), double buffered
 *                                              (only the grid-synchronised kernel uses it)
 *   tn (int32)           [P][T][A][J][D]       State::transnum, dp innermost
 *   tm (double)          [P][T][A][J][D]       State::transmean
 *   poptraj              [P][T][A][C][D]       State::pop (tests / outputs only)
 *   obs_partial          [P][nslot]            per (observation, area) running sums
 *   obs_val              [P][nobs]             ObservationModel::get_obs_value
 * A tile's J*D (or C*D) block is contiguous, so a warp reads and writes it with
 * unit-stride lanes.
 */
#ifndef BEEP_DEV_TYPES_H
#define BEEP_DEV_TYPES_H

#include <stdint.h>

struct ObsEnt { int ob; int slot; };

/* per-area constants of the mixing matrix, one 32-byte record (two 16-byte loads) */
struct alignas(16) AreaInfo { double diag, fac; int j0, nn, rg, ie0; };   /* ie0 = efoi_ref of strain 0 */
/* one entry of the transposed mixing matrix: weight and the source area's offset into a dinf plane */
struct alignas(16) Nbr { double val; unsigned off; int pad; };   /* off in BYTES: source area * G * 8 */

#define BEEP_EVEN(x) (((x) + 1) & ~1)

/* Shared-memory layout of the trajectory kernel, in doubles (every block even so that 16-byte
 * accesses stay aligned); computed identically by the launcher and by the kernel.
 *   CTA common : N0R[G][Gs] age matrix rows, stride Gs even with Gs/2 odd (conflict-free 16-byte row reads)
 *                ctab[JD32] packed cell table (u32), J*D rounded up to 32 with never-firing cells
 *   per warp   : ns[JD] ae[JD u16] foi[D] Iv[Gs] nbr[32 Nbr] nmi[G]      (work area)
 *   SYNC_NONE, per warp: pop[CDp] W dinf[2][S*G]
 *   SYNC_CTA,  per CTA:  pop[A][CDp] dinf[2][S*A*G]; per warp: W
 *   SYNC_GRID, per warp: pop[CDp] W                     (the tile streams in by cp.async while the
 *                                                        force of infection is computed) */
#ifndef BEEP_TRAJ_MINB
#define BEEP_TRAJ_MINB 4
#endif
#ifndef BEEP_TRAJ_NWARP
#define BEEP_TRAJ_NWARP 8                /* warps per CTA (upper bound; the grid flavour uses exactly this many) */
#endif
struct TrajSm { int Gs, n0, ctab, common, JDe, JD32, ns, ae, foi, nmi, iv, nbr, work; };
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline TrajSm traj_sm_layout(int G, int J, int D, int nbr_doubles = 64)
{
	TrajSm L;
	int gs = BEEP_EVEN(G);
	if (((gs >> 1) & 1) == 0) gs += 2;
	L.Gs = gs;
	L.n0 = 0; L.ctab = BEEP_EVEN(G * gs);
	L.JD32 = (J * D + 31) & ~31;                  /* cell table padded to whole warp passes */
	L.common = L.ctab + (L.JD32 >> 1);
	L.JDe = BEEP_EVEN(J * D);
	L.ns = 0; L.ae = L.JDe;
	L.foi = L.ae + BEEP_EVEN((L.JDe + 3) >> 2);
	L.iv = L.foi + BEEP_EVEN(D);
	L.nbr = L.iv + gs;
	L.nmi = L.nbr + BEEP_EVEN(nbr_doubles);   /* 32 staged Nbr entries; without neighbours only the staged spline values (if any) */
	L.work = L.nmi + BEEP_EVEN(G);
	return L;
}

/* Shared-memory layout of the chunk kernel (traj_chunk.cuh): byte offsets inside a warp's area.
 *   pop[NT][CDp4] int32 | lmean[cap] double (the infectivity vectors Iv[NT][Gs] during the force-of-infection
 *   phase) | foi[NT][G] double | tinfo[NT] 32-byte records | lcell[cap] u16 | acc[NT][CHUNK_NEJ] int32 | mbarrier |
 *   lu: 8 doubles + 8 ints */
#define CHUNK_NEJ 8                       /* observation entries per (sett, area) cell the chunk kernel accumulates */
struct ChunkSm { int pop, lmean, foi, tinfo, lcell, acc, bar, lu, per, cap; };
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline ChunkSm chunk_sm_layout(int NT, int G, int Gs, int JD32, int CDp4)
{
	ChunkSm s;
	int o = 0;
	s.cap = JD32 > 256 ? JD32 : 256;
	s.pop = o; o += NT * CDp4 * 4;
	s.lmean = o; o += (s.cap > NT * Gs ? s.cap : NT * Gs) * 8;
	s.foi = o; o += ((NT * G + 1) & ~1) * 8;
	s.tinfo = o; o += NT * 32;
	s.lcell = o; o += ((s.cap * 2 + 15) & ~15);
	s.acc = o; o += NT * CHUNK_NEJ * 4;
	s.bar = o; o += 16;
	s.lu = o; o += 96;                        /* (coefficient, byte offset) of up to 8 infectious compartments, staged for the close phase */
	s.per = (o + 15) & ~15;
	return s;
}

struct DevModel {
	int T, period, spu, A, C, J, D, G, S, Dq, nparam, nspline, nobs, ntable;
	int CDp4;           /* C*D rounded up to a multiple of 4: tile stride of the int32 populations (16-byte bulk copies) */
	unsigned obs_tr_any;            /* transitions that some observation counts */
	const unsigned char *cell_npop; /* [T*A] entries of the (sett, area) cell's list that observe a population; they come first */
	int stochastic, siminf, itr, geo_spline, nR, nefoi, ninf, nmm, ncovar, nlevel, ndir;
	int ndcc;                       /* demographic changes (data.democat_change, state.cc:339-376) */
	const int *dcc_info, *dcc_area, *dcc_group; const double *dcc_frac;   /* [ndcc][6] ncat, first area, areas, groups, first dp, first fraction */
	int nslot, nsection, has_region, has_area_eff, age_fast, M_nnz;
	int CDp;            /* C*D rounded up to even: tile stride of pop_cur (16-byte async copies) */
	int mm_packed;      /* every element of the age matrix is hit by at most three spline factors (and nspline < 64) */
	int obs_wide;       /* observations span many dp: reduce them warp-cooperatively */
	double popsize, dt;

	const int *dir_off, *dir_th; const double *dir_frac;
	const int *trans_from, *trans_to, *trans_inf; const double *trans_shape;
	const int *trans_mean_param, *trans_prob_param;
	const int *comp_inf_param, *inf_state, *inf_ref;
	const int *spline_factor_param, *spline_off, *spline_param; const double *spline_t, *spline_multfac;
	const int *sus_off, *sus_param;
	const int *region_effect_param, *area_effect_param, *covar_param; const double *covar_value;
	const int *level_param, *level_map; const double *area_total_pop;
	const double *N0, *N0T; const int *mm_spline, *mm_cnt; const int *dp_age;
	const unsigned *mm_packp;           /* the same words laid out like the padded shared-memory matrix, [G][Gs] (padding: all unused) */
	const unsigned *mm_pack;            /* [G*G] when mm_packed: up to three spline indices, one per byte (unused = nspline,
	                                       whose staged value is 1.0), so that an element is n0 * f0 * f1 * f2 exactly */
	const int *mm_prog_off, *mm_prog;   /* per element (a,b) of the age matrix: the splines that multiply it, in the
	                                       reference's order (model.cc:1254-1285; a spline appears once per hit) */
	const int *R_spline_ref, *area_rgrp; const double *R_dist; const int *strain_Rfactor_param;
	const int *efoi_spline_ref, *efoi_ref; const double *efoi_agedist;
	const int *Min_off, *Min_src; const double *Min_val; const double *M_diag, *M_factor;
	const double *Mt_dense;     /* [A][A] transposed mixing matrix, dense (zero diagonal), or null: the DMMA flavour of the contraction */
	const AreaInfo *area_info;  /* [A] */
	const Nbr *nbr;             /* [nnz] transposed mixing matrix, sources ascending per destination */
	const double *pop_init;
	const double *obs_value, *obs_sd, *obs_factor, *obs_shape;
	const int *obs_fs, *obs_table, *obs_model, *obs_section, *obs_slot_off;
	const unsigned *obs_tmask, *obs_cmask; const unsigned long long *obs_dmask;
	const int *cell_off; const ObsEnt *cell_ent;
	const unsigned *cell_tab;   /* [J*D] packed (pop offset | tr<<12 | dp<<18 | kind<<24) */
	const double *table_threshold;
	const int *section_ti, *section_tf;
	/* model modifications of mode=prediction (model.modelmod): null = none.  Models that carry one run the general tile kernel */
	int has_mod;
	const double *mod_tmean, *mod_beta, *mod_efoi, *mod_spline_mult, *mod_spline_set;
};

struct DevBatch {
	int P; unsigned flags;
	double *theta, *pd, *transrate, *disc_spline, *sus, *areafactor, *betaR, *infco;
	double *crate;                 /* [P][JD32] per-cell multiplier of dt*pop: transition rate, susceptibility for the
	                                  infection transition, 0 for cells that never fire (and the padding) */
	double *etab;                  /* [P][nefoi][T] external force of infection per individual, spline / popsize */
	int *inflist;                  /* [P][C+1] count, then the compartments with non-zero infectivity (ascending) */
	unsigned long long *sched;     /* [P][16] per particle, one 128-byte line: word 0 = (step << 32 | next tile) of the
	                                  particle-step scheduler, word 1 = tiles of that step completed */
	double *pop_cur, *dinf;
	double *imap;                  /* [P][S][A][G] Imap of the current step, computed by k_imap_dmma (dense mixing matrix only) */
	int *popi;                     /* [P][A][CDp4] int32 populations: batches the chunk kernel serves (then pop_cur is null) */
	double *scratchA;              /* [P][A] K0 scratch (per-area averages) */
	int *tn; double *tm; double *poptraj;
	double *obs_partial, *obs_val, *L, *EFtab, *Li;
	unsigned long long *stream_id;
	unsigned long long *infmask;   /* [P] transitions whose source or target compartment is infectious */
};

enum { MODE_SIM = 0, MODE_MBP = 1, MODE_REPLAY = 2 };
enum { SYNC_NONE = 0, SYNC_CTA = 1, SYNC_GRID = 2 };

struct TrajArgs {
	DevModel m;
	DevBatch b;
	int ti, tf;
	unsigned long long seed;
	unsigned rk[20];    /* the Philox round keys of `seed` (beep_philox4x32_10_rk): constant-bank operands of the draws */
	/* MBP: the coupled initial state (same slot of the `init` batch) */
	const int *init_tn; const double *init_tm; const unsigned char *simu_mask;
	const double *imap_pre;       /* non-null: Imap of this step for every (particle, strain, area, age), from k_imap_dmma */
	int nwarp;          /* warps per CTA */
	int chunk;          /* SYNC_GRID: tiles a warp takes per grab of the step's tile counter */
	int fast;           /* the model has the common structure tile_step_fast handles */
	float invA;         /* 1/A: tile index -> (particle, area) without an integer division */
	TrajSm L;           /* shared-memory layout (traj_sm_layout), computed once by the launcher */
	ChunkSm Ls;         /* chunk kernel: per-warp layout */
	unsigned invG;      /* chunk kernel: ceil(2^16 / G), pair index -> (tile, age) without a division */
	unsigned char tfrom[32], tto[32], tkind[32];   /* per transition: source, target, kind (0 ordinary, 1 infection, 2 never) */
};

#endif
