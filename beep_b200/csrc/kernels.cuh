/* kernels.cuh -- hand-written sm_100a kernels of the BEEP latent-state engine.
 *
 *   k_set_param, k_beta, k_areafactor_*
 *                 K0   State::set_param                    reference src/state.cc:66-91, src/model.cc:1010-1362
 *   k_traj<M,S,K> K1-3 State::simulate / Mbp::mbp /        src/state.cc:596-638,
 *                      State::initialise_from_particle     src/mbp.cc:252-330, src/state.cc:294-322
 *                      (+ obs accumulation, obsmodel.cc:74-109)
 *   k_obs_final   K4   ObservationModel::calculate*        src/obsmodel.cc:22-58,245-322
 *   k_latent_*    K5   State::likelihood                   src/state.cc:812-826
 *   k_gather_slab, k_copy_flagged
 *                 K6   Mpi::copy_particles / end_sec_swap  src/mpi.cc:81-165,215-327
 *   k_mvn_*       K7   ABCSMC::calculate_particle_weight   src/abcsmc.cc:260-271, src/mvn.cc:185-196
 *
 * Mapping of the trajectory kernel: ONE WARP PER (particle, area) TILE.  The
 * tile's population pop[C][D] lives in shared memory for the whole time loop
 * (SYNC_NONE / SYNC_CTA) or is streamed through shared memory once per step
 * (SYNC_GRID); lanes run over the J*D cells of the tile for the draws (one
 * Philox stream per cell) and over dp / age for the population update and the
 * force of infection.  The per-step coupling between areas (Imap needs dinf of
 * every source area, state.cc:204-216) is a barrier whose scope is chosen by
 * the launcher: none (uncoupled areas: persistent warps take trajectories from
 * a counter), __syncthreads (a CTA owns a particle), or one step barrier PER
 * PARTICLE kept in global memory (a particle spans the grid: warps take chunks
 * of tiles from the particle's scheduler word and move on to another particle
 * while a step's last tiles are in flight -- no grid-wide barrier).
 *
 * The translation unit is compiled with --fmad=false: every double operation
 * below is the same IEEE operation, in the same order, as in the CPU oracle
 * (oracle/beep_oracle.c, built -ffp-contract=off), so that transition means --
 * and hence the integer draws -- are bit-identical.
 */
#ifndef BEEP_KERNELS_CUH
#define BEEP_KERNELS_CUH

#include <cooperative_groups.h>
#include "dev_types.h"
#include "beep_b200.h"
#include "beep_rng.h"

namespace cg = cooperative_groups;

#define FULLMASK 0xffffffffu
#define BEEP_LARGE 1000000000.0

/* ------------------------------------------------------------------------ */
/* K0  set_param                                                             */
/* ------------------------------------------------------------------------ */

/* Ntime[t][a][b] = N0[a][b] * prod_mm f_mm^(cnt_mm[a]+cnt_mm[b])   (model.cc:1254-1285):
 * every matmod multiplies row a and column a of each listed age by its spline value. */
__device__ __forceinline__ double ntime_elem(const DevModel &m, const double *ds, int t, int a, int b, double n0)
{
	double x = n0;
	const int e = a * m.G + b;
	const double *dst = ds + t;
	for (int q = m.mm_prog_off[e], q1 = m.mm_prog_off[e + 1]; q < q1; q++) x *= dst[m.mm_prog[q] * m.T];
	return x;
}

/* true when no spline that modifies the age matrix changes value from step t-1 to step t:
 * Ntime[t] is then Ntime[t-1] bit for bit (lanes share the work; warp-uniform result) */
__device__ __forceinline__ bool ntime_same(const DevModel &m, const double *ds, int t, int lane)
{
	int ch = 0;
	for (int mm = lane; mm < m.nmm; mm += 32) { const double *q = ds + m.mm_spline[mm] * m.T + t; if (q[0] != q[-1]) ch = 1; }
	return !__any_sync(0xffffffffu, ch);
}

/* the step's spline values staged for the packed factor programs: fv[sp] = spline sp at step t, fv[nspline] = 1 */
__device__ __forceinline__ void ntime_stage(const DevModel &m, const double *ds, int t, double *fv, int lane)
{
	for (int sp = lane; sp <= m.nspline; sp += 32) fv[sp] = (sp < m.nspline) ? ds[sp * m.T + t] : 1.0;
	__syncwarp();
}
/* element e = a*G+b of Ntime from the staged values: n0 * f0 * f1 * f2 (multiplying by the 1.0 of an unused slot is exact) */
__device__ __forceinline__ double ntime_packed(const unsigned w, const double *fv, double n0)
{
	return n0 * fv[w & 255u] * fv[(w >> 8) & 255u] * fv[(w >> 16) & 255u];
}

/* the warp's copy of Ntime[t] (rows padded to Gs like n0r; the padding is zero), rebuilt from the
 * base matrix at the first step and wherever a modifying spline moved.  fv: nspline+1 doubles of
 * warp scratch (used when the model's factor programs are packed) */
__device__ __forceinline__ void ntime_refresh(const DevModel &m, const double *ds, int t, int t0, const double *n0r, double *nt, int Gs, double *fv, int lane)
{
	if (t != t0 && ntime_same(m, ds, t, lane)) return;
	__syncwarp();
	if (m.mm_packed) {
		/* branch-free over the padded layout: padding has n0 = 0 and multiplies by the staged 1.0 */
		ntime_stage(m, ds, t, fv, lane);
		const unsigned fa = (unsigned)__cvta_generic_to_shared(fv), na = (unsigned)__cvta_generic_to_shared(n0r), ta = (unsigned)__cvta_generic_to_shared(nt);
		const int n = m.G * Gs;
#pragma unroll 2
		for (int i = lane; i < n; i += 32) {
			const unsigned w = m.mm_packp[i];
			double x, f0, f1, f2;
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(na + i * 8));
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f0) : "r"(fa + ((w & 255u) << 3)));
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f1) : "r"(fa + ((w >> 5) & 0x7f8u)));
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f2) : "r"(fa + ((w >> 13) & 0x7f8u)));
			x = x * f0 * f1 * f2;
			asm volatile("st.shared.f64 [%0], %1;" ::"r"(ta + i * 8), "d"(x) : "memory");
		}
	} else {
		for (int i = lane; i < m.G * Gs; i += 32) {
			const int r = i / Gs, k = i - r * Gs;
			nt[i] = (k < m.G) ? ntime_elem(m, ds, t, r, k, n0r[i]) : 0.0;
		}
	}
	__syncwarp();
}

/* per-warp shared memory of k_beta in doubles: vinv0[K0_MAXINF][Dq] ngm[Dq][Dq+1] vec[Dq] vec2[Dq] ntc[G*G] fv[nspline+1] */
__host__ __device__ inline size_t k_beta_smem(int Dq, int G, int nspline, int nmm)
{
	return (size_t)8 * Dq + (size_t)Dq * (Dq + 1) + 2 * Dq + (nmm > 0 ? (size_t)G * G + nspline + 1 : 0);
}

#define K0_THREADS 128
#define K0_MAXINF 8
#define K0_BETA_WARPS 8

__global__ void __launch_bounds__(K0_THREADS) k_set_param(DevModel m, DevBatch b)
{
	const int p = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
	const int T = m.T, J = m.J, D = m.D;
	const double *theta = b.theta + (size_t)p * m.nparam;
	double *pd = b.pd + (size_t)p * m.nparam;
	double *transrate = b.transrate + (size_t)p * J * D;
	double *ds = b.disc_spline + (size_t)p * m.nspline * T;
	double *sus = b.sus + (size_t)p * D;
	double *infco = b.infco + (size_t)p * m.C;

	/* model.cc:2685-2695 dirichlet_correct */
	for (int i = tid; i < m.nparam; i += nth) pd[i] = theta[i];
	__syncthreads();
	for (int g = tid; g < m.ndir; g += nth) {
		double sum = 0;
		for (int k = m.dir_off[g]; k < m.dir_off[g + 1]; k++) sum += theta[m.dir_th[k]];
		for (int k = m.dir_off[g]; k < m.dir_off[g + 1]; k++) pd[m.dir_th[k]] = theta[m.dir_th[k]] / (m.dir_frac[k] * sum);
	}
	__syncthreads();

	/* model.cc:1116-1140 create_transrate */
	for (int e = tid; e < J * D; e += nth) {
		int tr = e / D;
		double rate = 0;
		if (!m.trans_inf[tr]) {
			double mean = pd[m.trans_mean_param[e]] / m.trans_shape[tr];
			rate = 1.0 / mean;
			if (m.trans_prob_param[e] >= 0) rate *= pd[m.trans_prob_param[e]];
		}
		transrate[e] = rate;
	}
	/* model.cc:1144-1164 create_susceptibility */
	for (int dp = tid; dp < D; dp += nth) {
		double val = 1.0;
		for (int k = m.sus_off[dp]; k < m.sus_off[dp + 1]; k++) val *= pd[m.sus_param[k]];
		sus[dp] = val;
	}
	for (int co = tid; co < m.C; co += nth) infco[co] = pd[m.comp_inf_param[co]];
	if (tid == 0) {
		/* transitions that can change the infectivity of a tile (either end infectious) */
		unsigned long long mk = 0;
		for (int tr = 0; tr < J; tr++) if (pd[m.comp_inf_param[m.trans_from[tr]]] != 0 || pd[m.comp_inf_param[m.trans_to[tr]]] != 0) mk |= 1ull << tr;
		b.infmask[p] = mk;
		/* the infectious compartments, ascending (the order state.cc:223-264 visits them in) */
		int *il = b.inflist + (size_t)p * (m.C + 1);
		int n = 0;
		for (int co = 0; co < m.C; co++) if (pd[m.comp_inf_param[co]] != 0) il[1 + n++] = co;
		il[0] = n;
	}
	/* model.cc:1020-1043 create_disc_spline, one thread per (spline, step): the knot interval is found
	 * by the reference's forward walk from the first knot; a step past the last knot repeats the value
	 * of the last step inside the knots (the reference's `prev`; 0 when there is none), filled in
	 * a second pass */
	for (int i = tid; i < m.nspline * T; i += nth) {
		const int sp = i / T, st = i - sp * T;
		const double factor = pd[m.spline_factor_param[sp]];
		const int o = m.spline_off[sp], np = m.spline_off[sp + 1] - o;
		const double t = (double)((st + 0.5) * m.period) / T;
		int q = 0;
		while (q + 1 < np && t > m.spline_t[o + q + 1]) q++;
		if (q + 1 < np) {
			const double fac = (t - m.spline_t[o + q]) / (m.spline_t[o + q + 1] - m.spline_t[o + q]);
			ds[i] = factor * (pd[m.spline_param[o + q]] * m.spline_multfac[o + q] * (1 - fac) +
			                  pd[m.spline_param[o + q + 1]] * m.spline_multfac[o + q + 1] * fac);
		}
	}
	__syncthreads();
	for (int i = tid; i < m.nspline * T; i += nth) {
		const int sp = i / T, st = i - sp * T;
		const int o = m.spline_off[sp], np = m.spline_off[sp + 1] - o;
		const double tl = m.spline_t[o + np - 1];
		if (np < 2 || (double)((st + 0.5) * m.period) / T > tl) {
			int sl = st - 1;                               /* last step at or before the last knot */
			while (sl >= 0 && (double)((sl + 0.5) * m.period) / T > tl) sl--;
			ds[i] = (sl >= 0 && np > 1) ? ds[sp * T + sl] : 0.0;
		}
	}
	__syncthreads();
	if (m.mod_spline_mult) {                                   /* model.cc:1044-1050: modifications of mode=prediction */
		for (int i = tid; i < m.nspline * T; i += nth) {
			ds[i] *= m.mod_spline_mult[i];
			const double set = m.mod_spline_set[i];
			if (set != 999999999.0) ds[i] = set;
		}
		__syncthreads();
	}
	/* per-cell multiplier table of the trajectory kernel's scan */
	{
		const int JD32 = (J * D + 31) & ~31;
		double *crate = b.crate + (size_t)p * JD32;
		for (int e = tid; e < JD32; e += nth) {
			double v = 0;
			if (e < J * D) { const int tr = e / D; v = (tr == m.itr) ? sus[e - tr * D] : transrate[e]; }
			crate[e] = v;
		}
	}
	/* external force of infection per individual (state.cc:423): spline value / population size */
	{
		double *etab = b.etab + (size_t)p * m.nefoi * T;
		for (int i = tid; i < m.nefoi * T; i += nth) { const int ie = i / T, t = i - ie * T; etab[i] = ds[m.efoi_spline_ref[ie] * T + t] / m.popsize; }
	}
}

/* model.cc:1318-1362 calculate_beta_from_R, ONE WARP per particle (K0_BETA_WARPS particles per CTA:
 * the eigenvalue iteration is a sequential chain per particle, so the machine is filled with
 * particles, not with threads of one particle).  The next-generation matrix NGM = F V^-1
 * (model.cc:2485-2576): V is block diagonal in dp, so only column (state 0, dp = i) of the
 * dp = i block of V^-1 contributes to NGM[q][i]; the blocks are inverted by the reference's
 * Gauss-Jordan (matrix.cc:703-765) and the terms are accumulated in the reference's order.
 * The dominant eigenvalue is the reference's power iteration (matrix.cc:1139-1220), started
 * from the flat vector and warm-started from solve to solve as the reference does -- every
 * reduction in the oracle's sequential order, so beta is bit-identical to the oracle's.
 * dynamic smem per warp: K0_MAXINF*Dq + Dq*(Dq+1) + 2*Dq doubles */
__global__ void __launch_bounds__(K0_BETA_WARPS * 32) k_beta(DevModel m, DevBatch b, int nwarp)
{
	extern __shared__ double sm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int p = blockIdx.x * nwarp + warp;
	if (p >= b.P) return;
	const int T = m.T, J = m.J, D = m.D, G = m.G, S = m.S, Dq = m.Dq;
	const double *pd = b.pd + (size_t)p * m.nparam;
	const double *transrate = b.transrate + (size_t)p * J * D;
	const double *ds = b.disc_spline + (size_t)p * m.nspline * T;
	const double *sus = b.sus + (size_t)p * D;
	double *betaR = b.betaR + (size_t)p * S * m.nR * T;
	const int ninf = m.ninf;
	double *vinv0 = sm + (size_t)warp * k_beta_smem(Dq, G, m.nspline, m.nmm);   /* [ninf][Dq] column 0 of each block inverse, per strain */
	double *ngm = vinv0 + K0_MAXINF * Dq, *vec = ngm + Dq * (Dq + 1), *vec2 = vec + Dq;
	double *ntc = vec2 + Dq;                             /* [G*G] Ntime of the last step that was evaluated */
	double *fv = ntc + G * G;                            /* [nspline+1] staged spline values */
	for (int i = lane; i < Dq; i += 32) vec[i] = 1.0 / Dq;
	for (int st = 0; st < S; st++) {
		__syncwarp();
		for (int dp = lane; dp < Dq; dp += 32) {
			double V[K0_MAXINF * K0_MAXINF], inv[K0_MAXINF * K0_MAXINF];
			for (int i = 0; i < ninf * ninf; i++) { V[i] = 0; inv[i] = 0; }
			for (int i = 0; i < ninf; i++) inv[i * ninf + i] = 1;
			for (int tr = 0; tr < J; tr++) {
				int from = m.inf_ref[m.trans_from[tr]], to = m.inf_ref[m.trans_to[tr]];
				if (from >= 0) {
					double rate = transrate[tr * D + st * Dq + dp];
					V[from * ninf + from] += rate;
					if (to >= 0) V[to * ninf + from] -= rate;
				}
			}
			for (int ii = 0; ii < ninf; ii++) {
				double r = V[ii * ninf + ii];
				for (int i = ii; i < ninf; i++) V[ii * ninf + i] /= r;
				for (int i = 0; i <= ii; i++) inv[ii * ninf + i] /= r;
				for (int jj = ii + 1; jj < ninf; jj++) {
					double rr = V[jj * ninf + ii];
					if (rr != 0) {
						for (int i = ii; i < ninf; i++) V[jj * ninf + i] -= rr * V[ii * ninf + i];
						for (int i = 0; i <= ii; i++) inv[jj * ninf + i] -= rr * inv[ii * ninf + i];
					}
				}
			}
			for (int ii = ninf - 1; ii > 0; ii--)
				for (int jj = ii - 1; jj >= 0; jj--) {
					double rr = V[jj * ninf + ii];
					if (rr != 0) for (int i = 0; i < ninf; i++) inv[jj * ninf + i] -= rr * inv[ii * ninf + i];
				}
			for (int is = 0; is < ninf; is++) vinv0[is * Dq + dp] = inv[is * ninf + 0];
		}
		__syncwarp();
		for (int ir = 0; ir < m.nR; ir++) {
			const int spl = m.R_spline_ref[ir];
			const double *dist = m.R_dist + (size_t)ir * Dq;
			const double rfac = pd[m.strain_Rfactor_param[st]];
			double *out = betaR + ((size_t)st * m.nR + ir) * T;
			__syncwarp();
			double ratio = 0;
			for (int t = 0; t < T; t++) {
				bool recompute = (t == 0);
				if (m.nmm > 0 && (t == 0 || !ntime_same(m, ds, t, lane))) {
					/* model.cc:1305-1314 equal(): any element of Ntime moved by more than TINY since the
					 * previous step (ntc holds Ntime[t-1]: steps skipped above left it bit-identical) */
					int ch = 0;
					if (m.mm_packed) ntime_stage(m, ds, t, fv, lane);
					for (int i = lane; i < G * G; i += 32) {
						double x;
						if (m.mm_packed) x = ntime_packed(m.mm_pack[i], fv, m.N0[i]);
						else { const int aa = i / G, bb = i - aa * G; x = ntime_elem(m, ds, t, aa, bb, m.N0[i]); }
						const double y = ntc[i];
						if (t > 0 && (y < x - 0.00000001 || y > x + 0.00000001)) ch = 1;
						ntc[i] = x;
					}
					if (__any_sync(FULLMASK, ch)) recompute = true;
					__syncwarp();
				}
				if (recompute) {
					for (int i = lane; i < Dq * Dq; i += 32) {
						int q = i / Dq, k = i - q * Dq;
						int aq = m.dp_age[q], ak = m.dp_age[k];
						double nt = (m.nmm > 0) ? ntc[aq * G + ak] : m.N0[aq * G + ak];
						double sum = 0;
						for (int is = 0; is < ninf; is++) {
							double inf = pd[m.comp_inf_param[m.inf_state[is]]];
							if (inf != 0) sum += dist[q] * sus[Dq * st + q] * inf * nt * vinv0[is * Dq + k];
						}
						ngm[q * (Dq + 1) + k] = sum;
					}
					__syncwarp();
					double ev = 0;
					for (unsigned loop = 0;; loop++) {
						for (int q = lane; q < Dq; q += 32) {
							double acc = 0;
							for (int k = 0; k < Dq; k++) acc += ngm[q * (Dq + 1) + k] * vec[k];
							vec2[q] = acc;
						}
						__syncwarp();
						double sum = 0;
						for (int k = 0; k < Dq; k++) sum += vec2[k];          /* every lane, the oracle's order */
						const double ev_new = sum;
						int bad = 0;
						double lim = 100 * 0.000000000000001;
						if (loop > 100) lim *= 10;
						if (loop > 1000) lim *= 10;
						__syncwarp();
						for (int q = lane; q < Dq; q += 32) {
							double v2 = vec2[q] / sum;
							double dif = v2 - vec[q];
							if (dif > lim || dif < -lim) bad = 1;
							vec2[q] = v2;
						}
						bool conv = false;
						if (loop % 10 == 0) {
							bool evok = (ev_new - ev > -lim && ev_new - ev < lim);
							conv = evok && !__any_sync(FULLMASK, bad);
						}
						__syncwarp();
						if (conv) break;                    /* as the reference: keep the pre-update ev and vec */
						for (int q = lane; q < Dq; q += 32) vec[q] = vec2[q];
						__syncwarp();
						ev = ev_new;
						if (loop + 1 > 10000) break;
					}
					ratio = ev;
				}
				if (m.nmm == 0) {
					/* age matrix constant in time: one eigenvalue serves every step */
					for (int tt = lane; tt < T; tt += 32) out[tt] = rfac * ds[spl * T + tt] / ratio;
					break;
				}
				if (lane == 0) out[t] = rfac * ds[spl * T + t] / ratio;
			}
			__syncwarp();
		}
	}
}

/* model.cc:1182-1250 create_areafactor + calculate_area_av, three launches so that large
 * area counts fill the machine.  exp() is the shared beep_exp and every sum runs in the
 * oracle's order (time average per area sequential in t, population-weighted sum sequential
 * in c), so the table is bit-identical to the CPU oracle's. */
__global__ void k_areafactor_fill(DevModel m, DevBatch b)
{
	const int p = blockIdx.x, c = blockIdx.y * blockDim.x + threadIdx.x, A = m.A, period = m.period;
	if (c >= A) return;
	const double *pd = b.pd + (size_t)p * m.nparam;
	double *af = b.areafactor + (size_t)p * period * A;
	double base = m.has_region ? beep_exp(pd[m.region_effect_param[c]]) : 1.0;
	if (m.has_area_eff) base *= pd[m.area_effect_param[c]];
	double av = 0;
	for (int t = 0; t < period; t++) {
		double v = base;
		for (int j = 0; j < m.ncovar; j++) v *= beep_exp(pd[m.covar_param[j]] * m.covar_value[((size_t)j * A + c) * period + t]);
		if (m.nlevel > 0) v *= pd[m.level_param[m.level_map[t * A + c]]];
		af[(size_t)t * A + c] = v;
		av += v;
	}
	av /= period;
	b.scratchA[(size_t)p * A + c] = av * m.area_total_pop[c];
}
__global__ void k_areafactor_sum(DevModel m, DevBatch b)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= b.P) return;
	const double *avs = b.scratchA + (size_t)p * m.A;
	double area_av = 0;
	for (int c = 0; c < m.A; c++) area_av += avs[c];
	b.Li[p] = area_av / m.popsize;          /* Li doubles as K0 scratch; it is rewritten by beep_latent_loglik */
}
__global__ void k_areafactor_scale(DevModel m, DevBatch b)
{
	const int p = blockIdx.x;
	const size_t n = (size_t)m.period * m.A;
	double *af = b.areafactor + (size_t)p * n;
	const double area_av = b.Li[p];
	for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) af[i] /= area_av;
}

/* ------------------------------------------------------------------------ */
/* K1-K3  trajectory kernel                                                  */
/* ------------------------------------------------------------------------ */

/* A tile in pop_cur is [C*D] doubles padded to an even count (stride CDp). */
__global__ void k_pop_reset(DevModel m, DevBatch b)
{
	const size_t CD = (size_t)m.C * m.D, CDp = m.CDp, n = (size_t)m.A * CDp, tot = n * b.P;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
		size_t r = i % n, c = r / CDp, k = r - c * CDp;
		b.pop_cur[i] = (k < CD) ? m.pop_init[c * CD + k] : 0.0;
	}
}

/* particle-step scheduler state: every particle stands at step ti with no tile handed out */
__global__ void k_sched_init(unsigned long long *sched, int P, int ti)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < P) { sched[(size_t)p * 16] = (unsigned long long)(unsigned)ti << 32; sched[(size_t)p * 16 + 1] = 0; }
	if (p == 0) sched[(size_t)P * 16] = 0;            /* particles that have closed their last step */
}

__global__ void k_zero(double *x, size_t n)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = 0;
}

/* 16-byte asynchronous global -> shared copies (LDGSTS), committed as one group per tile */
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
	unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

/* Explicit shared-memory accesses through 32-bit shared-window addresses.  The hot path keeps ONE
 * laundered base address per warp and adds layout offsets that sit in the constant bank; with C++
 * pointers the compiler re-derives every shared base (special registers + constants) at each use
 * once registers are tight, which cost ~15 % of the issued instructions (profiles/README.md). */
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned launder(unsigned x) { asm volatile("mov.u32 %0, %0;" : "+r"(x)); return x; }
__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ unsigned lds32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds16(unsigned a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ double2 lds128(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(unsigned a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }

/* gpu-scope primitives of the particle-step scheduler.  Mutable data that crosses warps (dinf, pop_cur,
 * the scheduler words) is only ever read at L2 (ld.relaxed.gpu / cp.async.cg), so publishing needs a
 * release (MEMBAR.GPU before the atomic) but no acquire fence -- which would invalidate the whole L1
 * (CCTL.IVALL) and with it the read-only model tables. */
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ld_l2(const double *p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ void red_add_release_u64(unsigned long long *p, unsigned long long x) { asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory"); }
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long x) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory"); }
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long x) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory"); }
/* dinf lives in global memory (grid flavour: read at L2) or in shared memory (generic pointer) */
template <bool DGLOB> __device__ __forceinline__ double ld_dinf(const double *p) { return DGLOB ? ld_l2(p) : *p; }

/* per-warp scratch, all in shared memory (layout: traj_sm_layout, dev_types.h) */
struct TileSm {
	double *pop;      /* [CDp] population of the tile                                   */
	double *ns;       /* [J*D] per cell: the mean of an active cell between scan and draw, then the
	                     event count of this step; ALL ZERO between tile steps          */
	double *foi;      /* [D]   force of infection per dp (state.cc:428)                */
	double *nmi;      /* [G]   NMI per age (state.cc:547-578)                           */
	double *Iv;       /* [Gs]  infectivity vector before the age matvec, zero padded   */
	unsigned short *ae; /* [J*D] cell index of the active cells, compacted             */
	Nbr *nbr;         /* [32]  staged entries of the transposed mixing matrix          */
	const double *N0R;     /* [G][Gs] age matrix, padded rows: the base matrix N0 (CTA shared), or the warp's
	                          own Ntime of the current step when it caches it (SYNC_NONE) */
	bool ntfly;            /* N0R is the base matrix and the model modifies it in time: apply
	                          model.cc:1254-1285 element by element                      */
	const unsigned *ctab;  /* [J*D] packed cell table (CTA shared)                      */
	int Gs;
};

/* what a warp needs to know about the tile it works on (all warp-uniform): the particle
 * part changes only when the warp moves to another particle, the area part per tile */
struct TileCtx {
	int p, c;
	const double *rate, *crate, *ds, *sus, *infco, *af, *betaR, *etab;
	const int *il;
	double *partial;
	unsigned long long sid;
	unsigned infmask;
	double diag, fac;
	int j0, nn, rg, ie0;
};

__device__ __forceinline__ void tile_ctx_particle(const TrajArgs &a, int p, TileCtx &tc)
{
	const DevModel &m = a.m;
	const DevBatch &b = a.b;
	tc.p = p;
	tc.rate = b.transrate + (size_t)p * (m.J * m.D);
	tc.crate = b.crate + (size_t)p * a.L.JD32;
	tc.ds = b.disc_spline + (size_t)p * (m.nspline * m.T);
	tc.sus = b.sus + (size_t)p * m.D;
	tc.infco = b.infco + (size_t)p * m.C;
	tc.il = b.inflist + (size_t)p * (m.C + 1);
	tc.af = b.areafactor + (size_t)p * m.period * m.A;
	tc.betaR = b.betaR + (size_t)p * (m.S * m.nR * m.T);
	tc.etab = b.etab + (size_t)p * (m.nefoi * m.T);
	tc.partial = b.obs_partial + (size_t)p * m.nslot;
	tc.sid = b.stream_id[p];
	tc.infmask = (unsigned)b.infmask[p];
}
__device__ __forceinline__ void tile_ctx_area(const AreaInfo &ai, int c, TileCtx &tc)
{
	tc.c = c;
	tc.diag = ai.diag; tc.fac = ai.fac; tc.j0 = ai.j0; tc.nn = ai.nn; tc.rg = ai.rg; tc.ie0 = ai.ie0;
}

/* dinf[st][a] of one tile from its population (state.cc:223-264): infectivity-weighted
 * number of individuals by strain and age, summed in the oracle's (co, dp) order over the
 * particle's infectious compartments (adding the exact zeros of empty cells, which the
 * reference skips, leaves every partial sum unchanged) */
__device__ __forceinline__ void tile_dinf(const DevModel &m, const double *infco, const int *il, const double *pop, double *out, int ostride, int lane)
{
	const int G = m.G, Dq = m.Dq, D = m.D;
	const int ninf = il[0];
	for (int st = 0; st < m.S; st++) {
		if (lane < G) {
			double sum = 0;
			for (int k = 1; k <= ninf; k++) {
				const int co = il[k];
				const double inf = infco[co];
				const double *pc = pop + co * D + st * Dq;
				if (m.age_fast) {
					for (int dp = lane; dp < Dq; dp += G) sum += inf * pc[dp];
				} else {
#pragma unroll 1
					for (int dp = 0; dp < Dq; dp++) if (m.dp_age[dp] == lane) sum += inf * pc[dp];
				}
			}
			out[st * ostride + lane] = sum;
		}
	}
}

/* One time step of one (particle, area) tile.  Returns the (warp-uniform) mask of
 * transitions that had events (0 at the last step, whose events are never applied).
 *   1. force of infection per dp;
 *   2. SCAN the J*D cells, 32 per pass: transition means; cells with a non-zero mean are
 *      compacted into a dense list (most cells of an epidemic are empty);
 *   3. DRAW on the dense list, 32 active cells per pass, one Philox stream per cell;
 *   4. observations, population update and infectivity, visiting only the
 *      transitions that had events.
 * KEEP = false compiles out every trajectory store (ABC-SMC / particle-filter use). */
template <int MODE, bool KEEP, bool LATEWAIT, bool DGLOB>
__device__ __forceinline__ unsigned tile_step(const TrajArgs &a, const TileCtx &tc, const int t, const int tday, const int lane, const TileSm &sm,
                                              const double *dinf_cur, double *dinf_nxt, const int Ad, const int ci)
{
	const DevModel &m = a.m;
	const DevBatch &b = a.b;
	const int T = m.T, A = m.A, J = m.J, D = m.D, G = m.G, S = m.S, Dq = m.Dq;
	const int JD = J * D;
	double *pop = sm.pop, *ns = sm.ns, *foi = sm.foi, *nmi = sm.nmi;
	const double *ds = tc.ds;

	/* LATEWAIT: the tile's population is still streaming into shared memory (cp.async); the force of
	 * infection does not read it, so the copy is awaited only after that phase */
	if (LATEWAIT && S > 1) { cp_async_wait<0>(); __syncwarp(); }
	/* state.cc:395-404: with several strains the susceptible pool lives in strain 0 */
	if (S > 1) {
		const int sfrom = a.tfrom[m.itr];
#pragma unroll 1
		for (int dp = lane; dp < Dq; dp += 32) {
			double sum = 0;
#pragma unroll 1
			for (int st = 1; st < S; st++) { sum += pop[sfrom * D + Dq * st + dp]; pop[sfrom * D + Dq * st + dp] = 0; }
			pop[sfrom * D + dp] += sum;
		}
		__syncwarp();
	}

	/* 1. state.cc:547-578 get_NMI + the infection part of set_transmean (:406-432) */
	{
		const double d = ds[m.geo_spline * T + t], omd = 1 - d;
		const double afv = tc.af[(size_t)tday * A + tc.c];
		const int ll = lane < G ? lane : G - 1;        /* idle lanes repeat the last age: every load stays in range */
		const bool fastdp = m.age_fast && Dq == G;     /* dp == age: the force of infection needs no NMI round trip */
		bool staged = false;
		for (int st = 0; st < S; st++) {
			const double *dc = dinf_cur + (size_t)st * Ad * G;
			const char *dcl = (const char *)(dc + ll);
			/* Imap of this area (state.cc:204-216 transposed): the entries of the area's column of M are
			 * staged through shared memory, 32 at a time, and read back as one 16-byte broadcast each;
			 * sources ascending = the reference's accumulation order */
			double Imap = 0;
			if (a.imap_pre) Imap = a.imap_pre[(((size_t)tc.p * S + st) * Ad + ci) * G + ll];      /* dense mixing matrix: one GEMM per step for the whole batch */
			else for (int jb = 0; jb < tc.nn; jb += 32) {
				const int cnt = (tc.nn - jb < 32) ? tc.nn - jb : 32;
				if (!staged) {
					__syncwarp();
					if (lane < cnt) sm.nbr[lane] = m.nbr[tc.j0 + jb + lane];
					__syncwarp();
					staged = (tc.nn <= 32);
				}
#pragma unroll 2
				for (int j = 0; j < cnt; j++) {
					const double2 raw = ((const double2 *)sm.nbr)[j];            /* one 16-byte broadcast: (weight, byte offset) */
					Imap += ld_dinf<DGLOB>((const double *)(dcl + (unsigned)__double2loint(raw.y))) * raw.x;
				}
			}
			const double Idiag = ld_dinf<DGLOB>(dc + ci * G + ll) * tc.diag;
			double nv = 0;                                 /* NMI of age `lane` */
			if (G == 1) {
				nv = Idiag * (d + tc.fac * omd) + Imap * d;
				if (nv < 0) nv = 0;
			} else {
				const double Iv = (Idiag + Imap) * d + tc.fac * Idiag * omd;
				/* a neighbourhood without infectious individuals has I = 0, hence NMI = 0 exactly */
				if (__any_sync(FULLMASK, Iv != 0)) {
					if (lane < G) sm.Iv[lane] = Iv;
					__syncwarp();
					if (lane < G) {
						double sum = 0;
						if (!sm.ntfly) {
							const double2 *row = (const double2 *)(sm.N0R + lane * sm.Gs);
							const double2 *iv2 = (const double2 *)sm.Iv;
							const int h = sm.Gs >> 1;
#pragma unroll 2
							for (int k = 0; k < h; k++) { const double2 n2 = row[k], v2 = iv2[k]; sum += n2.x * v2.x; sum += n2.y * v2.y; }
						} else {
#pragma unroll 1
							for (int aa = 0; aa < G; aa++) sum += ntime_elem(m, ds, t, lane, aa, sm.N0R[lane * sm.Gs + aa]) * sm.Iv[aa];
						}
						if (sum < 0) sum = 0;
						nv = sum;
					}
				}
			}
			double be = tc.betaR[(st * m.nR + tc.rg) * T + t];
			if (m.mod_beta) be *= m.mod_beta[((size_t)t * A + tc.c) * S + st];      /* model.cc:1349-1357 */
			be *= afv;
			const int ie = (st == 0) ? tc.ie0 : m.efoi_ref[st * A + tc.c];
			double et = tc.etab[ie * T + t];
			if (m.mod_efoi) et *= m.mod_efoi[((size_t)t * A + tc.c) * S + st];       /* state.cc:415 */
			if (fastdp) {
				if (lane < G) foi[st * Dq + lane] = be * nv + et * m.efoi_agedist[ie * G + lane];
			} else {
				if (lane < G) nmi[lane] = nv;
				__syncwarp();
#pragma unroll 1
				for (int dp = lane; dp < Dq; dp += 32) {
					const int ag = m.dp_age[dp];
					foi[st * Dq + dp] = be * nmi[ag] + et * m.efoi_agedist[ie * G + ag];
				}
			}
			__syncwarp();
		}
	}

	if (LATEWAIT && S == 1) { cp_async_wait<0>(); __syncwarp(); }

	/* 2. scan: transition means (state.cc:434-444); compact the cells that can have events */
	const size_t tbase = KEEP ? (((size_t)tc.p * T + t) * A + tc.c) * (size_t)JD : 0;
	const bool keep_tn = KEEP && (b.flags & BEEP_KEEP_TRANSNUM) != 0, keep_tm = KEEP && (b.flags & BEEP_KEEP_TRANSMEAN) != 0;
	const bool stoch = m.stochastic != 0;
	const unsigned lt = (1u << lane) - 1u;
	const double dt = m.dt;
	int nact = 0;
	unsigned trbits = 0;
#pragma unroll 2
	for (int e0 = 0; e0 < JD; e0 += 32) {
		const int e = e0 + lane;
		double mean = 0.0;
		bool act = false;
		unsigned w = 0;
		if (e < JD) {
			w = sm.ctab[e];
			const double popu = pop[w & 0xfffu];
			const double r = tc.rate[e];
			if (popu > 0) {
				const unsigned kind = w >> 24;
				if (kind == 0) mean = dt * popu * r;
				else if (kind == 1) { const int dp = (w >> 18) & 63; mean = dt * popu * tc.sus[dp] * foi[dp]; }
			}
			if (m.mod_tmean) mean *= m.mod_tmean[((size_t)t * A + tc.c) * JD + e];     /* state.cc:446-453 */
			if (MODE == MODE_REPLAY) {
				/* the events are given (state.cc:294-322); only the means are recomputed */
				const double num = stoch ? (double)b.tn[tbase + e] : b.tm[tbase + e];
				if (keep_tm) b.tm[tbase + e] = mean;
				if (num != 0) { ns[e] = num; trbits |= 1u << ((w >> 12) & 63); }
			} else {
				if (keep_tm) b.tm[tbase + e] = mean;
				if (!stoch) { if (mean != 0) { ns[e] = mean; trbits |= 1u << ((w >> 12) & 63); } }   /* state.cc:627 */
				else {
					act = (mean != 0);
					if (keep_tn && !act) b.tn[tbase + e] = 0;
				}
			}
		}
		if (MODE != MODE_REPLAY) {
			const unsigned bal = __ballot_sync(FULLMASK, act);
			if (act) { sm.ae[nact + __popc(bal & lt)] = (unsigned short)e; ns[e] = mean; }
			nact += __popc(bal);
		}
	}

	/* 3. draws on the dense list */
	if (MODE != MODE_REPLAY && nact > 0) {
		__syncwarp();
		const unsigned long long cell0 = ((unsigned long long)t * A + tc.c) * (unsigned long long)JD;
		int simu = 0;
		if (MODE == MODE_MBP && a.simu_mask) simu = a.simu_mask[((size_t)tc.p * T + t) * A + tc.c];
		for (int i = lane; i < nact; i += 32) {
			const int e = sm.ae[i];
			const double mean = ns[e];
			BeepStream rs;
			beep_stream_init_cell(&rs, a.seed, tc.sid, cell0 + e);
			rs.rk = a.rk;
			int num;
			if (MODE == MODE_SIM) num = beep_poisson(&rs, mean);             /* state.cc:624-626 */
			else {
				/* mbp.cc:286-311 */
				const int num_i = a.init_tn[tbase + e];
				if (simu) num = beep_poisson(&rs, mean);
				else {
					const double rate_i = a.init_tm[tbase + e];
					if (mean > rate_i) num = num_i + beep_poisson(&rs, mean - rate_i);
					else if (mean == rate_i) num = num_i;
					else if (num_i == 0) num = 0;
					else num = beep_binomial(&rs, mean / rate_i, num_i);
				}
			}
			ns[e] = (double)num;                            /* replaces the mean: zero again unless there were events */
			if (num != 0) trbits |= 1u << ((sm.ctab[e] >> 12) & 63);
			if (keep_tn) b.tn[tbase + e] = num;
		}
	}
	/* transitions with events, warp-wide (also orders the ns[] writes before the reads below) */
	const unsigned trmask = __reduce_or_sync(FULLMASK, trbits);

	/* 4a. obsmodel.cc:74-109: add this (sett, area) cell block to the observations that list it */
	{
		const int k0 = m.cell_off[t * A + tc.c], k1 = m.cell_off[t * A + tc.c + 1];
		if (m.obs_wide) {
			/* observations summing many dp: lanes over dp, one warp reduction per entry */
			for (int k = k0; k < k1; k++) {
				const ObsEnt en = m.cell_ent[k];
				const int ob = en.ob;
				const unsigned tmask0 = m.obs_tmask[ob] & trmask, cm0 = m.obs_cmask[ob];
				if (tmask0 == 0 && cm0 == 0) continue;         /* no events in the observed transitions */
				const unsigned long long dmask = m.obs_dmask[ob];
				const double factor = m.obs_factor[ob];
				const int fs = m.obs_fs[ob];
				const double dsv = (fs >= 0) ? ds[fs * T + t] : 1.0;
				double val = 0;
				for (int dp = lane; dp < D; dp += 32) {
					if (!((dmask >> dp) & 1ull)) continue;
					unsigned tmask = tmask0;
					while (tmask) {
						const int tr = __ffs(tmask) - 1; tmask &= tmask - 1;
						const double num = ns[tr * D + dp];
						if (num != 0) { if (fs < 0) val += factor * num; else val += factor * num * dsv; }
					}
					unsigned cm = cm0;
					while (cm) {
						const int co = __ffs(cm) - 1; cm &= cm - 1;
						val += factor * pop[co * D + dp];
					}
				}
				for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(FULLMASK, val, o);
				if (lane == 0 && val != 0) atomicAdd(&tc.partial[en.slot], val);
			}
		} else {
			/* narrow observations (one or a few dp each): one lane per entry */
			for (int k = k0 + lane; k < k1; k += 32) {
				const ObsEnt en = m.cell_ent[k];
				const int ob = en.ob;
				unsigned tmask = m.obs_tmask[ob] & trmask, cm = m.obs_cmask[ob];
				if (tmask == 0 && cm == 0) continue;
				const double factor = m.obs_factor[ob];
				const int fs = m.obs_fs[ob];
				const unsigned long long dmask = m.obs_dmask[ob];
				double val = 0;
				const double dsv = (fs >= 0) ? ds[fs * T + t] : 1.0;
				while (tmask) {
					const int tr = __ffs(tmask) - 1; tmask &= tmask - 1;
					unsigned long long dm = dmask;
					while (dm) {
						const int dp = __ffsll((long long)dm) - 1; dm &= dm - 1;
						const double num = ns[tr * D + dp];
						if (num != 0) { if (fs < 0) val += factor * num; else val += factor * num * dsv; }
					}
				}
				while (cm) {
					const int co = __ffs(cm) - 1; cm &= cm - 1;
					unsigned long long dm = dmask;
					while (dm) {
						const int dp = __ffsll((long long)dm) - 1; dm &= dm - 1;
						val += factor * pop[co * D + dp];
					}
				}
				/* the slot has a single owner (this tile) and steps are ordered: a plain RED */
				if (val != 0) atomicAdd(&tc.partial[en.slot], val);
			}
		}
	}

	if (KEEP && (b.flags & BEEP_KEEP_POP)) {
		double *pt = b.poptraj + (((size_t)tc.p * T + t) * A + tc.c) * (size_t)(m.C * D);
		for (int i = lane; i < m.C * D; i += 32) pt[i] = pop[i];
	}
	__syncwarp();

	/* 4b. state.cc:460-485 update_pop in transition order (the last step's events are never
	 * applied), consuming -- and clearing -- the event counts; then the infectivity of the
	 * new population.  Without events both are unchanged. */
	const bool apply = (t < T - 1);
	{
		unsigned tm_ = trmask;
		while (tm_) {
			const int tr = __ffs(tm_) - 1; tm_ &= tm_ - 1;
			const int fo = a.tfrom[tr] * D, to = a.tto[tr] * D;
			for (int dp = lane; dp < D; dp += 32) {
				const double num = ns[tr * D + dp];
				if (num != 0) {
					if (apply) { pop[fo + dp] -= num; pop[to + dp] += num; }
					ns[tr * D + dp] = 0;
				}
			}
		}
		__syncwarp();
		if (apply) {
			if (trmask & tc.infmask) tile_dinf(m, tc.infco, tc.il, pop, dinf_nxt + ci * G, Ad * G, lane);
			else if (lane < G)
#pragma unroll 1
				for (int st = 0; st < S; st++) dinf_nxt[(st * Ad + ci) * G + lane] = ld_dinf<DGLOB>(dinf_cur + (st * Ad + ci) * G + lane);
		}
	}
	__syncwarp();
	return apply ? trmask : 0u;
}

/* The same tile step for the common model structure -- one strain, dp == age (D == G > 1),
 * stochastic dynamics, MODE_SIM / MODE_MBP / MODE_REPLAY -- written against explicit shared addresses:
 *   wb    the warp's work area (ns, ae, foi, Iv, nbr), pa the tile's population, ca the CTA tables.
 * Same operations in the same order as tile_step (hence as the oracle); the scan is branch-free:
 * the cell table is padded to a multiple of 32 with never-firing cells and crate[] holds the
 * per-cell multiplier (rate, or susceptibility for the infection transition, 0 otherwise). */
template <int MODE, bool KEEP, bool LATEWAIT, bool DGLOB>
__device__ __forceinline__ unsigned tile_step_fast(const TrajArgs &a, const int p, const int c, const int t, const int tday, const int lane,
                                                   const unsigned wb, const unsigned pa, const unsigned ca,
                                                   double *dbase, const size_t dflip, const int ci, const unsigned n0_a, const bool ntfly)
{
	/* Loop-carried state is (p, c) and three shared addresses: every table pointer is re-derived from
	 * the kernel parameters where it is used, so that nothing else stays live across the phases. */
	const double *dinf_cur = dbase + (t & 1) * dflip;
	double *dinf_nxt = dbase + ((t + 1) & 1) * dflip;
	const DevModel &m = a.m;
	const DevBatch &b = a.b;
	const TrajSm &L = a.L;
	const int T = m.T, A = m.A, J = m.J, G = m.G;
	const int JD = J * G;
	const unsigned ns_a = wb + L.ns * 8, ae_a = wb + L.ae * 8, foi_a = wb + L.foi * 8, iv_a = wb + L.iv * 8, nbr_a = wb + L.nbr * 8;
	const unsigned ctab_a = ca + L.ctab * 8;
	const double *ds = b.disc_spline + (size_t)p * (m.nspline * T);

	/* 1. state.cc:547-578 get_NMI + the infection part of set_transmean (:406-432) */
	{
		const double d = ds[m.geo_spline * T + t], omd = 1 - d;
		const double afv = b.areafactor[((size_t)p * m.period + tday) * A + c];
		const AreaInfo tc = m.area_info[c];
		const int ll = lane < G ? lane : G - 1;        /* idle lanes repeat the last age: every load stays in range */
		const char *dcl = (const char *)(dinf_cur + ll);
		double Imap = 0;
		if (a.imap_pre) Imap = a.imap_pre[((size_t)p * A + c) * G + ll];           /* dense mixing matrix: one GEMM per step for the whole batch (S == 1 here) */
		else
#pragma unroll 1
		for (int jb = 0; jb < tc.nn; jb += 32) {
			const int cnt = (tc.nn - jb < 32) ? tc.nn - jb : 32;
			__syncwarp();
			if (lane < cnt) sts128(nbr_a + lane * 16, *(const double2 *)&m.nbr[tc.j0 + jb + lane]);
			__syncwarp();
#pragma unroll 5
			for (int j = 0; j < cnt; j++) {
				const double2 raw = lds128(nbr_a + j * 16);                  /* one 16-byte broadcast: (weight, byte offset) */
				Imap += ld_dinf<DGLOB>((const double *)(dcl + (unsigned)__double2loint(raw.y))) * raw.x;
			}
		}
		const double Idiag = ld_dinf<DGLOB>(dinf_cur + ci * G + ll) * tc.diag;
		const double Iv = (Idiag + Imap) * d + tc.fac * Idiag * omd;
		double nv = 0;
		/* a neighbourhood without infectious individuals has I = 0, hence NMI = 0 exactly */
		if (__any_sync(FULLMASK, Iv != 0)) {
			if (lane < G) sts64(iv_a + lane * 8, Iv);
			__syncwarp();
			if (lane < G) {
				double sum = 0;
				const unsigned row = n0_a + lane * (L.Gs * 8);
				if (!ntfly) {
					const int h = L.Gs >> 1;
#pragma unroll 4
					for (int k = 0; k < h; k++) { const double2 n2 = lds128(row + k * 16), v2 = lds128(iv_a + k * 16); sum += n2.x * v2.x; sum += n2.y * v2.y; }
				} else {
#pragma unroll 1
					for (int aa = 0; aa < G; aa++) sum += ntime_elem(m, ds, t, lane, aa, lds64(row + aa * 8)) * lds64(iv_a + aa * 8);
				}
				if (sum < 0) sum = 0;
				nv = sum;
			}
		}
		double be = b.betaR[((size_t)p * m.nR + tc.rg) * T + t];
		be *= afv;
		const double et = b.etab[((size_t)p * m.nefoi + tc.ie0) * T + t];
		if (lane < G) sts64(foi_a + lane * 8, be * nv + et * m.efoi_agedist[tc.ie0 * G + lane]);
		__syncwarp();
	}
	if (LATEWAIT) { cp_async_wait<0>(); __syncwarp(); }

	/* 2. scan: transition means (state.cc:434-444); compact the cells that can have events */
	const size_t tbase = (KEEP || MODE == MODE_MBP) ? (((size_t)p * T + t) * A + c) * (size_t)JD : 0;
	const bool keep_tn = KEEP && (b.flags & BEEP_KEEP_TRANSNUM) != 0, keep_tm = KEEP && (b.flags & BEEP_KEEP_TRANSMEAN) != 0;
	const unsigned lt = (1u << lane) - 1u;
	const double dt = m.dt;
	/* two lists in ae[]: cells with a small mean (Poisson by inversion) from the front, cells with a
	 * large mean (transformed rejection) from the back -- the two samplers never share a warp pass.
	 * A pass whose 32 cells all have an empty source compartment (most passes of an epidemic's
	 * early and late phases) leaves after four instructions. */
	unsigned cnt2 = 0;                                  /* nsmall | nbig << 16 (one register across the loop) */
	int simu = 0;
	if (MODE == MODE_MBP && a.simu_mask) simu = a.simu_mask[((size_t)p * T + t) * A + c];
	const int top = L.JDe - 1;
	unsigned trbits = 0;
	const double *crate = b.crate + (size_t)p * L.JD32;
#pragma unroll 1
	for (int e0 = 0; e0 < JD; e0 += 32) {
		const int e = e0 + lane;
		const unsigned w = lds32(ctab_a + e * 4);          /* padded to a multiple of 32 with never-firing cells */
		const double popu = lds64(pa + ((w & 0xfffu) << 3));
		if (MODE == MODE_REPLAY) {
			/* the events are given (state.cc:294-322); only the means are recomputed */
			if (e < JD) {
				double mean = dt * popu * crate[e];
				if ((w >> 24) == 1u) mean *= lds64(foi_a + ((w >> 15) & 0x1f8u));
				if (!(popu > 0)) mean = 0;
				const int num = b.tn[tbase + e];
				if (keep_tm) b.tm[tbase + e] = mean;
				if (num != 0) { sts64(ns_a + e * 8, (double)num); trbits |= 1u << ((w >> 12) & 63); }
			}
			continue;
		}
		if (!__any_sync(FULLMASK, popu > 0 && (w >> 25) == 0u)) {          /* kind 2 (and the padding) never fires */
			if (KEEP && e < JD) {
				if (keep_tm) b.tm[tbase + e] = 0;
				if (keep_tn) b.tn[tbase + e] = 0;
			}
			continue;
		}
		const double cr = crate[e];
		double mean = dt * popu * cr;
		if ((w >> 24) == 1u) mean *= lds64(foi_a + ((w >> 15) & 0x1f8u));
		if (!(popu > 0)) mean = 0;
		bool act = (mean != 0);
		if (KEEP && e < JD) {
			if (keep_tm) b.tm[tbase + e] = mean;
			if (keep_tn && !act) b.tn[tbase + e] = 0;
		}
		if (MODE == MODE_MBP && !DGLOB && act && !simu) {
			/* mbp.cc:286-311: cells whose proposal needs no random number are settled here, from the
			 * coalesced read of the initial state (which also brings its lines in for the draws):
			 * equal means keep the count, a smaller mean thins a zero count to zero.  (Not in the
			 * grid flavour, where the extra load latency in the scan costs more than it saves.) */
			const int num_i = a.init_tn[tbase + e];
			const double rate_i = a.init_tm[tbase + e];
			if (mean == rate_i || (mean < rate_i && num_i == 0)) {
				const int num = (mean == rate_i) ? num_i : 0;
				if (keep_tn) b.tn[tbase + e] = num;
				if (num != 0) { sts64(ns_a + e * 8, (double)num); trbits |= 1u << ((w >> 12) & 63); }
				act = false;
			}
		}
		const bool big = act && mean >= BEEP_POISSON_INV_MAX;
		const unsigned bal = __ballot_sync(FULLMASK, act), balb = __ballot_sync(FULLMASK, big);
		if (act) {
			const int pos = big ? top - (int)((cnt2 >> 16) + __popc(balb & lt)) : (int)((cnt2 & 0xffffu) + __popc(bal & ~balb & lt));
			sts16(ae_a + pos * 2, e); sts64(ns_a + e * 8, mean);
		}
		cnt2 += __popc(bal & ~balb) + (__popc(balb) << 16);
	}
	const int nsmall = (int)(cnt2 & 0xffffu), nbig = (int)(cnt2 >> 16);

	/* 3. draws on the dense lists */
	if (MODE != MODE_REPLAY && nsmall + nbig > 0) {
		__syncwarp();
		const unsigned long long cell0 = ((unsigned long long)t * A + c) * (unsigned long long)JD;
		const unsigned long long sid = b.stream_id[p];
		/* one cell: the draw of MODE (state.cc:624-626 / mbp.cc:286-311) from the cell's own stream */
		auto draw = [&](const int e, const double mean) -> int {
			BeepStream rs;
			beep_stream_init_cell(&rs, a.seed, sid, cell0 + e);
			rs.rk = a.rk;
			if (MODE == MODE_SIM) return beep_poisson(&rs, mean);
			const int num_i = a.init_tn[tbase + e];
			if (simu) return beep_poisson(&rs, mean);
			const double rate_i = a.init_tm[tbase + e];
			if (mean > rate_i) return num_i + beep_poisson(&rs, mean - rate_i);
			if (mean == rate_i) return num_i;
			if (num_i == 0) return 0;
			return beep_binomial(&rs, mean / rate_i, num_i);
		};
		/* replaces the mean by the count: zero again unless there were events */
		auto finish = [&](const int e, const int num) {
			sts64(ns_a + e * 8, (double)num);
			if (num != 0) trbits |= 1u << ((lds32(ctab_a + e * 4) >> 12) & 63);
			if (keep_tn) b.tn[tbase + e] = num;
		};
#pragma unroll 1
		for (int i = lane; i < nsmall; i += 32) {
			const int e = lds16(ae_a + i * 2);
			finish(e, draw(e, lds64(ns_a + e * 8)));
		}
		if (nbig > 0) {
			if (MODE == MODE_SIM) {
				/* first the squeeze of every cell's first attempt; the cells it does not settle are
				 * collected (over the slots already read) and run the full sampler together */
				int ndef = 0;
#pragma unroll 1
				for (int i0 = 0; i0 < nbig; i0 += 32) {
					const int i = i0 + lane;
					bool defer = false;
					int e = 0;
					if (i < nbig) {
						e = lds16(ae_a + (top - i) * 2);
						const double mean = lds64(ns_a + e * 8);
						BeepStream rs;
						beep_stream_init_cell(&rs, a.seed, sid, cell0 + e);
						int num;
						if (beep_poisson_ptrs_try(rs.key[0], rs.key[1], rs.ctr[0], rs.ctr[1], rs.ctr[2], rs.ctr[3], mean, &num, a.rk)) finish(e, num);
						else defer = true;
					}
					const unsigned bd = __ballot_sync(FULLMASK, defer);
					if (defer) sts16(ae_a + (top - (ndef + __popc(bd & lt))) * 2, e);
					ndef += __popc(bd);
				}
				__syncwarp();
#pragma unroll 1
				for (int i = lane; i < ndef; i += 32) {
					const int e = lds16(ae_a + (top - i) * 2);
					finish(e, draw(e, lds64(ns_a + e * 8)));
				}
			} else {
#pragma unroll 1
				for (int i = lane; i < nbig; i += 32) {
					const int e = lds16(ae_a + (top - i) * 2);
					finish(e, draw(e, lds64(ns_a + e * 8)));
				}
			}
		}
	}
	/* transitions with events, warp-wide (also orders the ns[] writes before the reads below) */
	const unsigned trmask = __reduce_or_sync(FULLMASK, trbits);

	/* 4a. obsmodel.cc:74-109: add this (sett, area) cell block to the observations that list it */
	{
		const int k0 = m.cell_off[t * A + c], k1 = m.cell_off[t * A + c + 1];
		double *partial = b.obs_partial + (size_t)p * m.nslot;
		if (m.obs_wide) {
			/* observations summing many dp: lanes over dp, one warp reduction per entry */
#pragma unroll 1
			for (int k = k0; k < k1; k++) {
				const ObsEnt en = m.cell_ent[k];
				const int ob = en.ob;
				const unsigned tmask0 = m.obs_tmask[ob] & trmask, cm0 = m.obs_cmask[ob];
				if (tmask0 == 0 && cm0 == 0) continue;         /* no events in the observed transitions */
				const unsigned long long dmask = m.obs_dmask[ob];
				const double factor = m.obs_factor[ob];
				const int fs = m.obs_fs[ob];
				const double dsv = (fs >= 0) ? ds[fs * T + t] : 1.0;
				double val = 0;
				if (lane < G && ((dmask >> lane) & 1ull)) {
					unsigned tmask = tmask0;
					while (tmask) {
						const int tr = __ffs(tmask) - 1; tmask &= tmask - 1;
						const double num = lds64(ns_a + (tr * G + lane) * 8);
						if (num != 0) { if (fs < 0) val += factor * num; else val += factor * num * dsv; }
					}
					unsigned cm = cm0;
					while (cm) {
						const int co = __ffs(cm) - 1; cm &= cm - 1;
						val += factor * lds64(pa + (co * G + lane) * 8);
					}
				}
				for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(FULLMASK, val, o);
				if (lane == 0 && val != 0) atomicAdd(&partial[en.slot], val);
			}
		} else {
			/* narrow observations (one or a few dp each): one lane per entry */
#pragma unroll 1
			for (int k = k0 + lane; k < k1; k += 32) {
				const ObsEnt en = m.cell_ent[k];
				const int ob = en.ob;
				unsigned tmask = m.obs_tmask[ob] & trmask, cm = m.obs_cmask[ob];
				if (tmask == 0 && cm == 0) continue;
				const double factor = m.obs_factor[ob];
				const int fs = m.obs_fs[ob];
				const unsigned long long dmask = m.obs_dmask[ob];
				double val = 0;
				const double dsv = (fs >= 0) ? ds[fs * T + t] : 1.0;
				while (tmask) {
					const int tr = __ffs(tmask) - 1; tmask &= tmask - 1;
					unsigned long long dm = dmask;
					while (dm) {
						const int dp = __ffsll((long long)dm) - 1; dm &= dm - 1;
						const double num = lds64(ns_a + (tr * G + dp) * 8);
						if (num != 0) { if (fs < 0) val += factor * num; else val += factor * num * dsv; }
					}
				}
				while (cm) {
					const int co = __ffs(cm) - 1; cm &= cm - 1;
					unsigned long long dm = dmask;
					while (dm) {
						const int dp = __ffsll((long long)dm) - 1; dm &= dm - 1;
						val += factor * lds64(pa + (co * G + dp) * 8);
					}
				}
				/* the slot has a single owner (this tile) and steps are ordered: a plain RED */
				if (val != 0) atomicAdd(&partial[en.slot], val);
			}
		}
	}
	__syncwarp();

	/* 4b. state.cc:460-485 update_pop in transition order (the last step's events are never
	 * applied), consuming -- and clearing -- the event counts; then the infectivity of the
	 * new population.  Without events both are unchanged. */
	const bool apply = (t < T - 1);
	{
		unsigned tm_ = trmask;
		while (tm_) {
			const int tr = __ffs(tm_) - 1; tm_ &= tm_ - 1;
			if (lane < G) {
				const unsigned na = ns_a + (tr * G + lane) * 8;
				const double num = lds64(na);
				if (num != 0) {
					if (apply) {
						const unsigned fa = pa + (a.tfrom[tr] * G + lane) * 8, ta = pa + (a.tto[tr] * G + lane) * 8;
						sts64(fa, lds64(fa) - num);
						sts64(ta, lds64(ta) + num);
					}
					sts64(na, 0.0);
				}
			}
		}
		__syncwarp();
		if (apply && lane < G) {
			double *out = dinf_nxt + ci * G + lane;
			if (trmask & (unsigned)b.infmask[p]) {
				/* state.cc:223-264: infectivity-weighted individuals by age over the infectious compartments */
				const int *il = b.inflist + (size_t)p * (m.C + 1);
				const double *infco = b.infco + (size_t)p * m.C;
				const int ninf = il[0];
				double sum = 0;
#pragma unroll 1
				for (int k = 1; k <= ninf; k++) { const int co = il[k]; sum += infco[co] * lds64(pa + (co * G + lane) * 8); }
				*out = sum;
			} else *out = ld_dinf<DGLOB>(dinf_cur + ci * G + lane);
		}
	}
	__syncwarp();
	return apply ? trmask : 0u;
}

/* CTA-shared tables at the front of dynamic shared memory */
/* State::democat_change_pop_adjust, state.cc:339-376, for one particle whose areas all sit in shared memory
 * (popA[c * CDp + co * D + dp]): the populations of a demographic change's areas are shifted between the categories
 * of its classification so that their fractions follow the data.  Same loops and the same ORDER of sums as the
 * reference (the oracle restates them): lane f accumulates category f over (area, group, compartment) in sequence,
 * the total is summed over f in sequence, then the (area, group, compartment) cells are shifted, `nth` threads of
 * the caller's group at a time.  scratch: 32 doubles of shared memory visible to the whole group.
 *   phase 1 is run by the group's first warp; sync() is the group's barrier (__syncwarp / __syncthreads). */
template <typename SyncFn>
__device__ __forceinline__ void democat_adjust(const DevModel &m, double *popA, const int CDp, const int t, const int tid, const int nth, double *scratch, SyncFn sync)
{
	const int D = m.D, C = m.C;
#pragma unroll 1
	for (int k = 0; k < m.ndcc; k++) {
		const int *inf = m.dcc_info + 6 * k;
		const int ncat = inf[0], na = inf[2], ng = inf[3];
		const int *area = m.dcc_area + inf[1], *grp = m.dcc_group + inf[4];
		if (tid < 32) {
			double pc = 0;
			if (tid < ncat)
				for (int ia = 0; ia < na; ia++) {
					const double *pa = popA + (size_t)area[ia] * CDp;
					for (int g = 0; g < ng; g++) { const int dp = grp[g * ncat + tid]; for (int co = 0; co < C; co++) pc += pa[co * D + dp]; }
				}
			double tot = 0;
			for (int f = 0; f < ncat; f++) tot += __shfl_sync(FULLMASK, pc, f);
			if (tid < ncat) scratch[tid] = m.dcc_frac[inf[5] + (size_t)t * ncat + tid] - pc / tot;
		}
		sync();
		const int nitem = na * ng * C;
		for (int it = tid; it < nitem; it += nth) {
			const int ia = it / (ng * C), r = it - ia * (ng * C), g = r / C, co = r - g * C;
			double *pa = popA + (size_t)area[ia] * CDp + co * D;
			double sum = 0;
			for (int f = 0; f < ncat; f++) sum += pa[grp[g * ncat + f]];
			for (int f = 0; f < ncat; f++) pa[grp[g * ncat + f]] += scratch[f] * sum;
		}
		sync();
	}
}

__device__ __forceinline__ double *sm_common_init(const DevModel &m, const TrajSm &L, double *smem, TileSm &ts)
{
	double *N0R = smem + L.n0;
	unsigned *ctab = (unsigned *)(smem + L.ctab);
	for (int i = threadIdx.x; i < m.G * L.Gs; i += blockDim.x) { const int r = i / L.Gs, k = i - r * L.Gs; N0R[i] = (k < m.G) ? m.N0[r * m.G + k] : 0.0; }
	for (int i = threadIdx.x; i < L.JD32; i += blockDim.x) ctab[i] = (i < m.J * m.D) ? m.cell_tab[i] : (2u << 24);
	ts.N0R = N0R; ts.ctab = ctab; ts.Gs = L.Gs;
	return smem + L.common;
}

__device__ __forceinline__ void work_init(TileSm &ts, const TrajSm &L, double *w, int lane)
{
	ts.ns = w + L.ns; ts.ae = (unsigned short *)(w + L.ae);
	ts.foi = w + L.foi; ts.nmi = w + L.nmi; ts.Iv = w + L.iv; ts.nbr = (Nbr *)(w + L.nbr);
	for (int i = lane; i < L.JDe; i += 32) ts.ns[i] = 0;
	for (int i = lane; i < L.Gs; i += 32) ts.Iv[i] = 0;
	__syncwarp();
}

/* per-warp pointers of the generic tile step, rebuilt where it is called (cold path) */
__device__ __forceinline__ void work_ptrs(TileSm &ts, const TrajSm &L, double *w)
{
	ts.ns = w + L.ns; ts.ae = (unsigned short *)(w + L.ae);
	ts.foi = w + L.foi; ts.nmi = w + L.nmi; ts.Iv = w + L.iv; ts.nbr = (Nbr *)(w + L.nbr);
}

template <int MODE, int SYNC, bool KEEP>
__global__ void __launch_bounds__(BEEP_TRAJ_NWARP * 32, BEEP_TRAJ_MINB) k_traj(const __grid_constant__ TrajArgs a)
{
	extern __shared__ __align__(16) double smraw[];
	const DevModel &m = a.m;
	const DevBatch &b = a.b;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = a.nwarp;
	const int A = m.A, G = m.G, S = m.S;
	const int CDp = m.CDp;
	const TrajSm &L = a.L;
	{
		TileSm ts0;
		sm_common_init(m, L, smraw, ts0);
	}
	__syncthreads();
	const unsigned ca = launder(smem_addr(smraw));
	const bool fast = a.fast != 0;
	double *sm = smraw + L.common;

	if (SYNC == SYNC_NONE) {
		/* a warp keeps its tile for the whole trajectory, so it also keeps Ntime: the modified age
		 * matrix is rebuilt only at steps where a modifying spline changes value */
		const int ntsz = (m.nmm > 0) ? G * L.Gs : 0;
		const int per = CDp + L.work + BEEP_EVEN(2 * S * G) + ntsz;
		double *pop = sm + (size_t)warp * per;
		double *ntw = ntsz ? pop + CDp + L.work + BEEP_EVEN(2 * S * G) : nullptr;
		const unsigned n0_a = launder(smem_addr(ntw ? ntw : smraw + L.n0));
		{
			TileSm ts0;
			work_init(ts0, L, pop + CDp, lane);
		}
		double *dinf = pop + CDp + L.work;
		/* Trajectories differ a lot in cost (an epidemic that takes off draws from the rejection samplers for
		 * hundreds of steps, one that dies out touches almost nothing), so tiles are not bound to warps: a
		 * persistent warp takes the next tile from a counter when it has finished one -- no CTA waits for its
		 * slowest warp and no wave waits for its slowest CTA. */
		const long long ntile = (long long)b.P * A;
#pragma unroll 1
		while (true) {
		long long tile = 0;
		if (lane == 0) tile = (long long)atomicAdd(b.sched, 1ull);
		tile = __shfl_sync(FULLMASK, tile, 0);
		if (tile >= ntile) break;
		const int p = (int)(tile / A), c = (int)(tile % A);
		double *gp = b.pop_cur + (size_t)tile * CDp;
		for (int i = lane; i < CDp; i += 32) pop[i] = gp[i];
		__syncwarp();
		tile_dinf(m, b.infco + (size_t)p * m.C, b.inflist + (size_t)p * (m.C + 1), pop, dinf + (a.ti & 1) * S * G, G, lane);
		__syncwarp();
		if (fast) {
			const unsigned wb = launder(smem_addr(pop + CDp)), pa = launder(smem_addr(pop));
#pragma unroll 1
			for (int t = a.ti; t < a.tf; t++) {
				if (ntw) ntime_refresh(m, b.disc_spline + (size_t)p * (m.nspline * m.T), t, a.ti, smraw + L.n0, ntw, L.Gs, pop + CDp + L.nbr, lane);
				tile_step_fast<MODE, KEEP, false, false>(a, p, c, t, t / m.spu, lane, wb, pa, ca, dinf, (size_t)(S * G), 0, n0_a, false);
			}
		} else {
			TileSm ts;
			ts.N0R = ntw ? ntw : smraw + L.n0; ts.ntfly = false; ts.ctab = (const unsigned *)(smraw + L.ctab); ts.Gs = L.Gs; ts.pop = pop;
			work_ptrs(ts, L, pop + CDp);
			TileCtx tc;
			tile_ctx_particle(a, p, tc);
			tile_ctx_area(m.area_info[c], c, tc);
#pragma unroll 1
			for (int t = a.ti; t < a.tf; t++) {
				if (m.ndcc) {
					/* state.cc:607 / mbp.cc:265 / state.cc:308 (A == 1 here: the launcher puts several areas in one CTA) */
					democat_adjust(m, pop, CDp, t, lane, 32, ts.ns, [] { __syncwarp(); });
					for (int i = lane; i < 32 && i < L.JDe; i += 32) ts.ns[i] = 0;
					__syncwarp();
				}
				if (ntw) ntime_refresh(m, tc.ds, t, a.ti, smraw + L.n0, ntw, L.Gs, pop + CDp + L.nbr, lane);
				tile_step<MODE, KEEP, false, false>(a, tc, t, t / m.spu, lane, ts, dinf + (t & 1) * S * G, dinf + ((t + 1) & 1) * S * G, 1, 0);
			}
		}
		for (int i = lane; i < CDp; i += 32) gp[i] = pop[i];
		__syncwarp();
		}
	} else if (SYNC == SYNC_CTA) {
		const int p = blockIdx.x;
		double *popA = sm, *dinf = popA + (size_t)A * CDp;
		double *wk = dinf + BEEP_EVEN(2 * S * A * G) + (size_t)warp * L.work;
		{
			TileSm ts0;
			work_init(ts0, L, wk, lane);
		}
		double *gp = b.pop_cur + (size_t)p * A * CDp;
		for (int i = threadIdx.x; i < A * CDp; i += blockDim.x) popA[i] = gp[i];
		__syncthreads();
		for (int c = warp; c < A; c += nwarp) tile_dinf(m, b.infco + (size_t)p * m.C, b.inflist + (size_t)p * (m.C + 1), popA + (size_t)c * CDp, dinf + (a.ti & 1) * S * A * G + c * G, A * G, lane);
		__syncthreads();
		const unsigned wb = launder(smem_addr(wk)), pa0 = launder(smem_addr(popA));
#pragma unroll 1
		for (int t = a.ti; t < a.tf; t++) {
			const int tday = t / m.spu;
			if (m.ndcc) {
				/* state.cc:607 / mbp.cc:265 / state.cc:308: one reduction over the particle's areas, then every thread shifts */
				double *scr = dinf + BEEP_EVEN(2 * S * A * G) + L.ns;      /* warp 0's mean buffer: idle between steps */
				democat_adjust(m, popA, CDp, t, (int)threadIdx.x, (int)blockDim.x, scr, [] { __syncthreads(); });
				if (threadIdx.x < 32 && (int)threadIdx.x < L.JDe) scr[threadIdx.x] = 0;
				__syncthreads();
			}
#pragma unroll 1
			for (int c = warp; c < A; c += nwarp) {
				if (fast) tile_step_fast<MODE, KEEP, false, false>(a, p, c, t, tday, lane, wb, pa0 + c * (CDp * 8), ca, dinf, (size_t)(S * A * G), c, ca + L.n0 * 8, m.nmm != 0);
				else {
					TileSm ts;
					ts.N0R = smraw + L.n0; ts.ntfly = m.nmm != 0; ts.ctab = (const unsigned *)(smraw + L.ctab); ts.Gs = L.Gs; ts.pop = popA + (size_t)c * CDp;
					work_ptrs(ts, L, wk);
					TileCtx tc;
					tile_ctx_particle(a, p, tc);
					tile_ctx_area(m.area_info[c], c, tc);
					tile_step<MODE, KEEP, false, false>(a, tc, t, tday, lane, ts, dinf + (t & 1) * S * A * G, dinf + ((t + 1) & 1) * S * A * G, A, c);
				}
			}
			__syncthreads();
		}
		for (int i = threadIdx.x; i < A * CDp; i += blockDim.x) gp[i] = popA[i];
	} else {
		cg::grid_group grid = cg::this_grid();
		const int per = CDp + L.work;
		double *popbuf = sm + (size_t)warp * per;
		{
			TileSm ts0;
			work_init(ts0, L, popbuf + CDp, lane);
		}
		const unsigned wb = launder(smem_addr(popbuf + CDp)), pa0 = launder(smem_addr(popbuf));
		const int ntile = b.P * A, nw = gridDim.x * nwarp, gw = blockIdx.x * nwarp + warp;   /* 32-bit: the launcher checks P*A < 2^30 */
		const size_t dsz = (size_t)b.P * S * A * G, psz = (size_t)S * A * G;
		const int nchunk = CDp >> 1;                    /* 16-byte chunks per tile */
		const int CH = a.chunk;
		/* infectivity of the starting population (static tile split: done once) */
#pragma unroll 1
		for (int tile = gw; tile < ntile; tile += nw) {
			const int p = tile / A, c = tile - p * A;
			const double *gp = b.pop_cur + (size_t)tile * CDp;
#pragma unroll 1
			for (int i = lane; i < CDp; i += 32) popbuf[i] = gp[i];
			__syncwarp();
			tile_dinf(m, b.infco + (size_t)p * m.C, b.inflist + (size_t)p * (m.C + 1), popbuf, b.dinf + (a.ti & 1) * dsz + (size_t)p * psz + c * G, A * G, lane);
			__syncwarp();
		}
		grid.sync();
		/* PARTICLE-STEP SCHEDULER.  Areas of one particle couple from step to step (Imap of step t+1
		 * needs dinf of every source area after step t), particles never do.  So instead of a grid
		 * barrier per step, each particle carries its own step barrier: sched[p] word 0 holds
		 * (step << 32 | next tile); a warp grabs CH consecutive tiles of the particle it is on
		 * with one 64-bit atomicAdd, works them, and adds them to word 1 (tiles completed) with
		 * release semantics; the warp that completes the step's last tile resets word 1 and publishes
		 * ((step + 1) << 32 | 0).  A warp that finds its particle exhausted (the step's last tiles
		 * are in flight) moves on to the next particle, whose step is in full swing -- the tail of
		 * one particle-step overlaps the head of another, and no warp waits for the grid.
		 * The tile's population streams into shared memory by cp.async (awaited after the
		 * force-of-infection phase, which does not read it) and is written back once per step. */
		const int P = b.P;
		int p = (int)(((long long)gw * P) / nw);        /* warps start spread over the particles */
		/* a warp that finds its particle exhausted moves on by its OWN stride (coprime with P, so P hops
		 * still visit every particle): warps leaving a particle together scatter instead of arriving
		 * together at the next one and exhausting it at once */
		int hop = 1;
		if (P > 2) {
			hop = 1 + (int)(((unsigned)gw * 2654435761u >> 8) % (unsigned)(P - 1));
			while (true) {
				int x = hop, y = P;
				while (y) { const int r = x % y; x = y; y = r; }
				if (x == 1) break;
				if (++hop >= P) hop = 1;
			}
		}
		int nmiss = 0;                                  /* consecutive particles without a tile to hand out */
		int nfin = 0;                                   /* ... of which had finished their trajectory */
		int nspin = 0;                                  /* idle rounds in a row: a scheduler that never hands out work traps */
		unsigned long long v = 0;                       /* lane 0: the grab in hand */
		bool have = false;                              /* v was grabbed ahead, during the previous chunk */
#pragma unroll 1
		while (true) {
			unsigned long long *sp = b.sched + (size_t)p * 16;
			if (!have) {
				if (lane == 0) {
					v = ld_relaxed_u64(sp);
					if ((int)(v >> 32) < a.tf && (unsigned)v < (unsigned)A) v = atomicAdd(sp, (unsigned long long)CH);
				}
			}
			have = false;
			v = __shfl_sync(FULLMASK, v, 0);
			const int t = (int)(v >> 32);
			const unsigned idx = (unsigned)v;
			if (t >= a.tf || idx >= (unsigned)A) {
				/* nothing to take here: finished, or the step's last tiles are still in flight */
				nmiss++;
				if (t >= a.tf) nfin++;
				/* every particle has closed its last step (counted by the closers): nothing is left anywhere */
				if ((nmiss & 3) == 0 && ld_relaxed_u64(b.sched + (size_t)P * 16) >= (unsigned long long)P) break;
				if (nmiss >= P) {
					if (nfin >= P) break;                   /* a full round of finished particles */
					nmiss = 0; nfin = 0;
					if (++nspin > 8000000) __trap();
					__nanosleep(500);
				}
				p += hop;
				if (p >= P) p -= P;
				continue;
			}
			nmiss = 0; nfin = 0; nspin = 0;
			const int c0 = (int)idx, c1 = (c0 + CH < A) ? c0 + CH : A;
			{
				const double *gp = b.pop_cur + ((size_t)p * A + c0) * CDp;
#pragma unroll 1
				for (int i = lane; i < nchunk; i += 32) cp_async16(popbuf + 2 * i, gp + 2 * i);
				cp_async_commit();
			}
			/* the next grab of this particle travels while this chunk is worked (a grab past the
			 * step's last tile is a miss and only overshoots the tile field, reset with the step) */
			unsigned long long vn = 0;
			if (c1 < A) { have = true; if (lane == 0) vn = atomicAdd(sp, (unsigned long long)CH); }
			const int tday = t / m.spu;
#pragma unroll 1
			for (int c = c0; c < c1; c++) {
				unsigned ev;
				if (fast) ev = tile_step_fast<MODE, KEEP, true, true>(a, p, c, t, tday, lane, wb, pa0, ca, b.dinf + (size_t)p * psz, dsz, c, ca + L.n0 * 8, m.nmm != 0);
				else {
					TileSm ts;
					ts.N0R = smraw + L.n0; ts.ntfly = m.nmm != 0; ts.ctab = (const unsigned *)(smraw + L.ctab); ts.Gs = L.Gs; ts.pop = popbuf;
					work_ptrs(ts, L, popbuf + CDp);
					TileCtx tc;
					tile_ctx_particle(a, p, tc);
					tile_ctx_area(m.area_info[c], c, tc);
					ev = tile_step<MODE, KEEP, true, true>(a, tc, t, tday, lane, ts, b.dinf + (t & 1) * dsz + (size_t)p * psz, b.dinf + ((t + 1) & 1) * dsz + (size_t)p * psz, A, c);
				}
				if (ev != 0 || S > 1) {
					const double2 *s2 = (const double2 *)popbuf;
					double2 *g2 = (double2 *)(b.pop_cur + ((size_t)p * A + c) * CDp);
#pragma unroll 1
					for (int i = lane; i < nchunk; i += 32) g2[i] = s2[i];
				}
				__syncwarp();
				if (c + 1 < c1) {
					const double *gn = b.pop_cur + ((size_t)p * A + c + 1) * CDp;
#pragma unroll 1
					for (int i = lane; i < nchunk; i += 32) cp_async16(popbuf + 2 * i, gn + 2 * i);
					cp_async_commit();
				}
			}
			/* publish: the lanes' stores are ordered before lane 0's release by the warp barrier.
			 * Completions are fire-and-forget REDs; the warp that took the step's LAST chunk closes the
			 * step: it waits for the chunks still in flight (their warps never wait for anything, so
			 * this ends), clears the counter and opens the next step. */
			__syncwarp();
			if (lane == 0) {
				red_add_release_u64(sp + 1, (unsigned long long)(c1 - c0));
				if (c1 == A) {
					int spin = 0;
					while (ld_relaxed_u64(sp + 1) != (unsigned long long)A) { if (++spin > 20000000) __trap(); __nanosleep(64); }
					st_relaxed_u64(sp + 1, 0ull);
					st_release_u64(sp, (unsigned long long)(unsigned)(t + 1) << 32);
					if (t + 1 >= a.tf) red_add_release_u64(b.sched + (size_t)P * 16, 1ull);
				}
			}
			v = vn;
		}
	}
}

/* ------------------------------------------------------------------------ */
/* K4  observation model                                                     */
/* ------------------------------------------------------------------------ */
__device__ __forceinline__ double d_normal_probability(double x, double mean, double var)
{
	return -0.5 * log(2 * 3.14159265358979323846 * var) - (x - mean) * (x - mean) / (2 * var);   /* utils.cc:107-111 */
}
__device__ __forceinline__ double d_poisson_probability(int i, double lam)                       /* utils.cc:157-170 */
{
	if (lam == 0) return (i == 0) ? 0.0 : -BEEP_LARGE;
	if (i < 0) return -BEEP_LARGE;
	return i * log(lam) - lam - lgamma((double)(i + 1));
}
__device__ __forceinline__ double d_negbin_probability(unsigned val, double mm, double shape)    /* utils.cc:124-127 */
{
	return lgamma(shape + val) - lgamma((double)(val + 1)) - lgamma(shape) + shape * log(shape / (shape + mm)) + val * log(mm / (shape + mm));
}

/* obsmodel.cc:245-305 obs_prob */
__device__ double d_obs_prob(const DevModel &m, double value, int ob)
{
	if (value < 0) value = 0;
	const double val = m.obs_value[ob], sd = m.obs_sd[ob];
	if (val == BEEP_UNKNOWN) return 0;
	const int om = m.obs_model[ob];
	if (val == BEEP_THRESH) {
		const double thresh = m.table_threshold[m.obs_table[ob]];
		if (om == BEEP_OBS_NORMAL || om == BEEP_OBS_NORMAL_PERCENT) {
			if (value < thresh) return d_normal_probability(thresh, thresh, sd * sd);
			return d_normal_probability(thresh, value, sd * sd);
		}
		if (om == BEEP_OBS_POISSON) {
			double lam = value; if (lam < 1) lam = 1;
			double sum = 0; for (double v = 0; v <= thresh; v++) sum += exp(d_poisson_probability((int)v, lam));
			return log(sum);
		}
		double mm = value; if (mm < 1) mm = 1;
		double sum = 0; for (double v = 0; v <= thresh; v++) sum += exp(d_negbin_probability((unsigned)v, mm, m.obs_shape[ob]));
		return log(sum);
	}
	if (om == BEEP_OBS_NORMAL || om == BEEP_OBS_NORMAL_PERCENT) return d_normal_probability(val, value, sd * sd);
	if (om == BEEP_OBS_POISSON) { double lam = value; if (lam < 1) lam = 1; return d_poisson_probability((int)(val + 0.5), lam); }
	double mm = value; if (mm < 1) mm = 1;
	return d_negbin_probability((unsigned)(int)(val + 0.5), mm, m.obs_shape[ob]);
}

#define RED_THREADS 256
/* deterministic block sum (fixed tree), result valid in thread 0 */
__device__ __forceinline__ double block_sum(double v, double *red)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0;
	if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
	return s;
}

/* obs_value, log-probability, L and the per-table split.  The observations of a particle are cut
 * into OBS_CHUNKS fixed chunks (one CTA each, fixed reduction tree), summed in chunk order by
 * k_obs_reduce -- deterministic, and enough CTAs for a million observations per particle.
 * section < 0: every observation (obsmodel.cc:22-35); else that section only (:39-58).
 * part is [P][OBS_CHUNKS][1 + ntable]. */
#define OBS_CHUNKS 32
__global__ void __launch_bounds__(RED_THREADS) k_obs_final(DevModel m, DevBatch b, int section, int want_tables, double *part)
{
	__shared__ double red[RED_THREADS / 32];
	const int p = blockIdx.x, ch = blockIdx.y, nch = gridDim.y;      /* nch <= OBS_CHUNKS: as many chunks as the observations fill */
	const int lo = (int)((long long)m.nobs * ch / nch), hi = (int)((long long)m.nobs * (ch + 1) / nch);
	const double *pt = b.obs_partial + (size_t)p * m.nslot;
	double *oval = b.obs_val + (size_t)p * m.nobs;
	double *out = part + ((size_t)p * OBS_CHUNKS + ch) * (1 + m.ntable);
	double acc = 0;
	for (int ob = lo + threadIdx.x; ob < hi; ob += blockDim.x) {
		if (section >= 0 && m.obs_section[ob] != section) continue;
		double v = 0;
		for (int k = m.obs_slot_off[ob]; k < m.obs_slot_off[ob + 1]; k++) v += pt[k];
		oval[ob] = v;
		acc += d_obs_prob(m, v, ob);
	}
	double L = block_sum(acc, red);
	if (threadIdx.x == 0) out[0] = L;
	if (want_tables) {
		for (int tb = 0; tb < m.ntable; tb++) {
			double a2 = 0;
			for (int ob = lo + threadIdx.x; ob < hi; ob += blockDim.x)
				if (m.obs_table[ob] == tb && (section < 0 || m.obs_section[ob] == section)) a2 += -2 * d_obs_prob(m, oval[ob], ob);
			double s = block_sum(a2, red);
			if (threadIdx.x == 0) out[1 + tb] = s;
		}
	}
}
__global__ void k_obs_reduce(DevModel m, int P, int want_tables, int nch, const double *part, double *L_out, double *EF_out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int nq = want_tables ? 1 + m.ntable : 1;
	if (i >= P * nq) return;
	const int p = i / nq, q = i - p * nq;
	double s = 0;
	for (int ch = 0; ch < nch; ch++) s += part[((size_t)p * OBS_CHUNKS + ch) * (1 + m.ntable) + q];
	if (q == 0) L_out[p] = s;
	else EF_out[(size_t)p * m.ntable + (q - 1)] = s;
}

/* ------------------------------------------------------------------------ */
/* K5  latent-process likelihood, state.cc:812-826                           */
/* ------------------------------------------------------------------------ */
/* The cells of a particle are cut into nch chunks (one CTA each, fixed reduction tree) that k_latent_final
 * adds in chunk order.  nch depends on the model only (k5_chunks), never on the batch, so a particle's
 * likelihood does not depend on the particles around it.  Most cells of an epidemic have no event:
 * k = 0 gives exactly -lam (0 log(lam) = 0, lgamma(1) = 0) and k = 1 needs no lgamma -- the kernel then
 * runs at the speed of its two input streams (4 + 8 bytes per cell). */
#define K5_MAX_CHUNKS 1024
static inline int k5_chunks(size_t ncell) { return (int)std::max<size_t>(1, std::min<size_t>(K5_MAX_CHUNKS, ncell / 65536)); }
/* log(k!) for k < 32 (correctly rounded; lgamma(k+1) agrees to an ulp or two -- the contract on L is 1e-10) */
__device__ const double d_lfact32[32] = {0.0, 0.0, 0.693147180559945, 1.7917594692280554, 3.178053830347945, 4.787491742782047, 6.579251212010102, 8.525161361065415, 10.604602902745249, 12.801827480081467, 15.104412573075514, 17.502307845873887, 19.987214495661885, 22.55216385312342, 25.191221182738683, 27.89927138384089, 30.671860106080672, 33.50507345013689, 36.39544520803305, 39.339884187199495, 42.335616460753485, 45.38013889847691, 48.47118135183522, 51.60667556776438, 54.78472939811232, 58.00360522298052, 61.26170176100201, 64.55753862700634, 67.88974313718153, 71.257038967168, 74.65823634883017, 78.0922235533153};
__device__ __forceinline__ double d_poisson_probability_fast(int i, double lam)                  /* utils.cc:157-170 */
{
	if (lam == 0) return (i == 0) ? 0.0 : -BEEP_LARGE;
	if (i < 0) return -BEEP_LARGE;
	if (i == 0) return -lam;                            /* 0 log(lam) - lam - lgamma(1) with lam in (0, inf): exactly -lam */
	if (i == 1) return log(lam) - lam;                  /* lgamma(2) = 0 */
	return i * log(lam) - lam - (i < 32 ? d_lfact32[i] : lgamma((double)(i + 1)));
}
__global__ void __launch_bounds__(RED_THREADS) k_latent_partial(DevModel m, DevBatch b, double *partial)
{
	__shared__ double red[RED_THREADS / 32];
	const int p = blockIdx.x, ch = blockIdx.y, nch = gridDim.y;
	const size_t n = (size_t)m.T * m.A * m.J * m.D;
	const size_t lo = n / nch * ch + min((size_t)ch, n % nch), hi = lo + n / nch + ((size_t)ch < n % nch ? 1 : 0);
	const int *tn = b.tn + (size_t)p * n;
	const double *tm = b.tm + (size_t)p * n;
	double acc = 0;
	for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
		double lam = tm[i];
		int k = m.stochastic ? tn[i] : (int)lam;
		acc += d_poisson_probability_fast(k, lam);
	}
	double s = block_sum(acc, red);
	if (threadIdx.x == 0) partial[(size_t)p * nch + ch] = s;
}
__global__ void k_latent_final(int P, int nch, const double *partial, double *Li)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < P) { double s = 0; for (int i = 0; i < nch; i++) s += partial[(size_t)p * nch + i]; Li[p] = s; }
}

/* ------------------------------------------------------------------------ */
/* K6  particle movement                                                     */
/* ------------------------------------------------------------------------ */
/* dst[p] <- src[map[p]] for slabs of `n` elements; map[p] == 0xFFFFFFFF keeps.  Out-of-place
 * (src != dst) so that chains of copies are well defined, as in mpi.cc:81-165. */
template <typename Tp>
__global__ void k_gather_slab(Tp *dst, const Tp *src, const unsigned *map, size_t n, int P)
{
	const int p = blockIdx.x;
	unsigned s = map ? map[p] : (unsigned)p;
	if (s == 0xFFFFFFFFu) s = (unsigned)p;
	const Tp *sp = src + (size_t)s * n;
	Tp *dp = dst + (size_t)p * n;
	for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) dp[i] = sp[i];
}
/* dst[map[k]] <- src[k]: the compact scratch of the moved slots written into place */
template <typename Tp>
__global__ void k_scatter_slab(Tp *dst, const Tp *src, const unsigned *map, size_t n)
{
	const Tp *sp = src + (size_t)blockIdx.x * n;
	Tp *dp = dst + (size_t)map[blockIdx.x] * n;
	for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) dp[i] = sp[i];
}
/* the same for a LIST of per-particle tables at once (the derived tables of K0 move with a particle: a dozen small slabs):
 * block (k, s, chunk) copies a piece of slab s of moved slot k, as 32-bit words */
struct SlabList { void *ptr[24]; unsigned words[24]; unsigned long long off[25]; int n; };
__global__ void k_gather_multi(const SlabList L, unsigned *tmp, const unsigned *map, int nm)
{
	const int k = blockIdx.x, s = blockIdx.y;
	const unsigned w = L.words[s];
	const unsigned *sp = (const unsigned *)L.ptr[s] + (size_t)map[k] * w;
	unsigned *dp = tmp + L.off[s] * (unsigned long long)nm + (size_t)k * w;
	for (size_t i = blockIdx.z * (size_t)blockDim.x + threadIdx.x; i < w; i += (size_t)gridDim.z * blockDim.x) dp[i] = sp[i];
}
__global__ void k_scatter_multi(const SlabList L, const unsigned *tmp, const unsigned *map, int nm)
{
	const int k = blockIdx.x, s = blockIdx.y;
	const unsigned w = L.words[s];
	const unsigned *sp = tmp + L.off[s] * (unsigned long long)nm + (size_t)k * w;
	unsigned *dp = (unsigned *)L.ptr[s] + (size_t)map[k] * w;
	for (size_t i = blockIdx.z * (size_t)blockDim.x + threadIdx.x; i < w; i += (size_t)gridDim.z * blockDim.x) dp[i] = sp[i];
}
/* dst[p] <- src[p] where flag[p] != 0 */
template <typename Tp>
__global__ void k_copy_flagged(Tp *dst, const Tp *src, const unsigned char *flag, size_t n, int P)
{
	const int p = blockIdx.x;
	if (!flag[p]) return;
	const Tp *sp = src + (size_t)p * n;
	Tp *dp = dst + (size_t)p * n;
	for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) dp[i] = sp[i];
}

/* ------------------------------------------------------------------------ */
/* K7  ABC-SMC importance-weight denominators, abcsmc.cc:260-271             */
/* ------------------------------------------------------------------------ */
/* The reference evaluates, for every proposal i, SUM_j w_j exp(-1/2 (d/size)' inv_M (d/size)) with
 * d = prop_i - last_j over the free parameters (MVN::probability, mvn.cc:185-196): O(N Nlast nvar^2).
 * With inv_M / size^2 = Lm Lm' (Cholesky, host, extended precision) the quadratic form is
 * |Lm'(x - y)|^2: both sets are whitened once (k_mvn_whiten, O((N + Nlast) nvar^2)) around the
 * weighted mean of the last generation (keeps magnitudes at the scale of d), and the pair sums
 * cost O(N Nlast nvar) (k_mvn_pair_sums).  Agrees with the reference's own sum to ~1e-12 relative
 * at cond(M) = 5e5 (tests/test_mvn_*.py); the contract is 1e-10. */
__global__ void k_mvn_whiten(const double *param, int n, int nparam, const int *var, int nvar, const double *mu, const double *Lm, double *out)
{
	/* out[i][k] = SUM_v (param[i][var[v]] - mu[v]) Lm[v][k],  k >= v only (Lm lower triangular), v ascending */
	const int i = blockIdx.x;
	const double *x = param + (size_t)i * nparam;
	for (int k = threadIdx.x; k < nvar; k += blockDim.x) {
		double acc = 0;
		for (int v = k; v < nvar; v++) acc = fma(x[var[v]] - mu[v], Lm[(size_t)v * nvar + k], acc);
		out[(size_t)i * nvar + k] = acc;
	}
}

#define MVN_TI 64
#define MVN_TJ 64
#define MVN_KC 16
/* partial[js][i] = SUM_{j in slice js} w[j] exp(-1/2 |xw_i - yw_j|^2).  CTA = 64 proposals x one slice of
 * the last generation, walked in tiles of 64; 256 threads, each a 4 x 4 block of pairs; operands staged
 * k-major in shared memory.  Every sum has a fixed order (k ascending; j by tile, then the 16 threads
 * of a row in lane order), so the result does not depend on the launch. */
__global__ void __launch_bounds__(256) k_mvn_pair_sums(const double *xw, int N, const double *yw, const double *w, int Nl, int nvar, int jslice, double *partial)
{
	__shared__ __align__(16) double xs[MVN_KC][MVN_TI + 2], ys[MVN_KC][MVN_TJ + 2];
	__shared__ double red[MVN_TI][17];
	const int i0 = blockIdx.x * MVN_TI, js = blockIdx.y;
	const int jbeg = js * jslice, jend = min(Nl, jbeg + jslice);
	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          /* tx: 4 columns (j), ty: 4 rows (i) */
	double rowsum[4] = {0, 0, 0, 0};
	for (int j0 = jbeg; j0 < jend; j0 += MVN_TJ) {
		double acc[4][4];
#pragma unroll
		for (int a = 0; a < 4; a++)
#pragma unroll
			for (int b = 0; b < 4; b++) acc[a][b] = 0;
		for (int k0 = 0; k0 < nvar; k0 += MVN_KC) {
			__syncthreads();
			for (int e = threadIdx.x; e < MVN_KC * MVN_TI; e += 256) {
				const int r = e / MVN_KC, k = e - r * MVN_KC;               /* consecutive threads read consecutive k: coalesced rows */
				const int i = i0 + r, j = j0 + r, kk = k0 + k;
				xs[k][r] = (i < N && kk < nvar) ? xw[(size_t)i * nvar + kk] : 0.0;
				ys[k][r] = (j < jend && kk < nvar) ? yw[(size_t)j * nvar + kk] : 0.0;
			}
			__syncthreads();
#pragma unroll
			for (int k = 0; k < MVN_KC; k++) {
				double xv[4], yv[4];
#pragma unroll
				for (int a = 0; a < 4; a++) { xv[a] = xs[k][ty * 4 + a]; yv[a] = ys[k][tx * 4 + a]; }
#pragma unroll
				for (int a = 0; a < 4; a++)
#pragma unroll
					for (int b = 0; b < 4; b++) { const double d = xv[a] - yv[b]; acc[a][b] = fma(d, d, acc[a][b]); }
			}
		}
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const int j = j0 + tx * 4 + b;
			if (j < jend) {
				const double wj = w[j];
#pragma unroll
				for (int a = 0; a < 4; a++) rowsum[a] += wj * exp(-0.5 * acc[a][b]);
			}
		}
	}
	__syncthreads();
#pragma unroll
	for (int a = 0; a < 4; a++) red[ty * 4 + a][tx] = rowsum[a];
	__syncthreads();
	if (threadIdx.x < MVN_TI && i0 + threadIdx.x < N) {
		double s = 0;
		for (int t = 0; t < 16; t++) s += red[threadIdx.x][t];
		partial[(size_t)js * N + i0 + threadIdx.x] = s;
	}
}
__global__ void k_mvn_reduce(const double *partial, int N, int nslice, double *out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double s = 0;
	for (int k = 0; k < nslice; k++) s += partial[(size_t)k * N + i];
	out[i] = s;
}

/* sampler parity helpers (tests/test_rng_parity.py) */
__global__ void k_poisson_many(unsigned long long seed, unsigned long long stream_id, const double *lam, int n, int *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { BeepStream rs; beep_stream_init(&rs, seed, (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)i); out[i] = beep_poisson(&rs, lam[i]); }
}
__global__ void k_binomial_many(unsigned long long seed, unsigned long long stream_id, const double *pr, const int *nn, int n, int *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { BeepStream rs; beep_stream_init(&rs, seed, (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)i); out[i] = beep_binomial(&rs, pr[i], nn[i]); }
}

#endif
