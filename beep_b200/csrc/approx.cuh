/* approx.cuh -- the MVN-approximate likelihood of mode=map (row N3 of SURVEY.md 8f), batched over particles.
 *
 *   k_likelihood_approx      State::likelihood_approx         reference src/state_mvn_approx.cc:11-91
 *     ap_step_means          set_Imap_sett / get_NMI / set_transmean   state.cc:164-290, 547-578, 380-456
 *     ap_covar_R, _update    set_transmean_gradients + covar_change_fast   state_mvn_approx.cc:1199-1284, 1081-1111
 *                            (covar_matrix_mult_SIMD / covar_add_mat, matrix.cc:199-236)
 *     ap_valid               State::valid_matrix              state_mvn_approx.cc:1133-1155
 *     ap_observation         State::apply_observation_fast    state_mvn_approx.cc:688-817
 *     ap_invert_det          invert_determinant_SIMD          matrix.cc:18-62, 564-579
 *
 * The reference carries, for ONE parameter vector at a time, the mean population p (n = areas x compartments x
 * demographic groups values) and its n x n covariance through the T steps: per step the transition means, their
 * sparse Jacobian J (transitions x compartments), covar += diag-like(mu) + S + S^T with S = Tm (J covar) (Tm the
 * +-1 incidence matrix), a validity clamp; at every observation time a Kalman-type update through two LU
 * inversions.  ML::generate_samples (ml.cc:389-413) evaluates it for every sample of a CMA-ES generation, and the
 * Hessian code (ml.cc:540-575) for O(nvar^2) parameter vectors: that loop is the batch dimension here.
 *
 * Mapping: ONE CTA PER PARTICLE (one to eight warps, by the size of the covariance), the whole time loop inside
 * the kernel (no launch per step).  The step vectors, and the covariance with the product R = J covar when they
 * leave room for several CTAs per SM, live in shared memory; otherwise (and for the LU factors and the mean
 * populations of all steps) in a per-particle scratch in HBM that the CTA re-reads through L1/L2.  Every phase is a
 * block-wide pass over (row, column) outputs -- a row per group of lanes, unit-stride columns, no index divisions --
 * separated by __syncthreads.  The Jacobian's dense part (the infection rows: one entry per source area, infectious
 * compartment and group) is rebuilt every step as (column, value) lists in ascending compartment order = the
 * reference's order of sums; the other rows have a single entry.  Sums that the reference does sequentially are
 * sequential here too (one thread per output), so results agree with the CPU restatement to rounding.
 *
 * As in the reference: one strain; as many transitions as non-susceptible compartments (state_mvn_approx.cc:1376);
 * no observation on the susceptible compartment.  The infectivity map is recomputed from the mean population
 * (see oracle/beep_oracle_approx.c for why that equals the reference's incremental map).
 */
#ifndef BEEP_APPROX_CUH
#define BEEP_APPROX_CUH

#include "kernels.cuh"

#define AP_THREADS 256
#define AP_LARGE 1000000000.0
#define AP_TINY 0.00000001

struct DevApprox {
	int n, nt, nwo, nws, start, nslice;
	const int *inc_off, *inc_ta;              /* per compartment ca: its transitions (ta ascending)            */
	const int *ta_from, *ta_to;               /* [nt] compartment indices                                      */
	const int *wo, *ws, *wsl_off, *wsl;       /* comp_without_S, comp_with_S, comp_with_S_list                 */
	const double *total_pop;                  /* [A][D]                                                        */
	const int *ocl_off, *ocl, *ocw_off, *ocw; /* obs_clist, obs_clist_without_S                                */
	const double *var_approx;                 /* [nobs] data.cc:1085-1115                                      */
	const int *obs_type, *obs_sett_i;         /* [nobs] table type (consts.hh:84), first step of the datum     */
	const int *sl_sett, *sl_off, *sl_obs;     /* observation slices, obsmodel.cc:357-385                       */
	/* per-particle scratch */
	double *pop;                              /* [P][T+1][n] mean populations                                  */
	double *covar, *R;                        /* [P][n][n], [P][nt][n]                                         */
	double *lu;                               /* [P][6][nwo][nwo]: M (in/out), l, u, linv, uinv, second matrix */
	double *vec;                              /* [P][AP_NVEC][max(n,nt)]                                       */
	double *L;                                /* [P]                                                           */
	double *jval; int *jcol;                  /* [P][jcap] the infection rows of the Jacobian, rebuilt every step      */
	long long jcap;
	int logW;                                 /* lanes per matrix row in the (row, column) passes: 2^logW, at most 32 */
	int sm_vec, sm_mat;                       /* the step vectors / covar and R live in dynamic shared memory          */
};
#define AP_NVEC 12
enum { AV_DINF = 0, AV_IV, AV_NMI, AV_BE, AV_MU, AV_DS, AV_SIG, AV_MU1, AV_MINVMU, AV_MUN, AV_VEC, AV_TMP };

/* block-wide sequential sum by thread 0, broadcast */
__device__ __forceinline__ double ap_bcast(double v, double *slot)
{
	__syncthreads();
	if (threadIdx.x == 0) *slot = v;
	__syncthreads();
	const double r = *slot;
	return r;
}

/* matrix.cc:18-62 + 564-579: Minv = M^-1 by Crout LU without pivoting (a zero pivot divides as TINY), det = sum log|l_ii|.
 * Right-looking form of the same elimination: at step i the row of U is scaled and the trailing block loses the
 * outer product l_:i u_i: -- every element sees its products one fused multiply-add at a time in ascending i, the
 * sequence of the CPU restatement -- so all threads work at every step.  L (with its diagonal) and U (unit diagonal)
 * share W; the inverses of the two triangular factors (the reference's row recurrence, invert_lower_triangular)
 * advance together, row i updating all rows below it.  M, Minv, W, li, ui: nw x nw; fv: nw.  Called by every thread. */
__device__ void ap_invert_det(const double *M, double *Minv, double *Wk, double *fv, double *li, double *ui, const int nw, const int logW, double *det_out, double *sm_slot)
{
	const int tid = threadIdx.x, nth = blockDim.x;
	const int W = 1 << logW, col0 = tid & (W - 1), row0 = tid >> logW, rstep = nth >> logW;
	for (int e = tid; e < nw * nw; e += nth) Wk[e] = M[e];
	__syncthreads();
	for (int i = 0; i < nw; i++) {
		double val = Wk[i * nw + i];
		if (val == 0) val = AP_TINY;
		for (int k = i + 1 + tid; k < nw; k += nth) Wk[i * nw + k] = Wk[i * nw + k] / val;
		__syncthreads();
		for (int j = i + 1 + row0; j < nw; j += rstep) {
			const double lji = -Wk[j * nw + i];
			for (int k = i + 1 + col0; k < nw; k += W) Wk[j * nw + k] = fma(lji, Wk[i * nw + k], Wk[j * nw + k]);
		}
		__syncthreads();
	}
	/* det = sum_i log(|u_ii| |l_ii|), u_ii = 1 (matrix.cc:44-48): the logarithms in parallel, the sum in index order */
	for (int i = tid; i < nw; i += nth) {
		double vall = Wk[i * nw + i];
		fv[i] = 1.0 / vall;
		if (vall < 0) vall = -vall;
		Minv[i] = log(1.0 * vall);
	}
	for (int e = tid; e < nw * nw; e += nth) { li[e] = 0; ui[e] = 0; }
	__syncthreads();
	if (tid == 0) {
		double d = 0;
		for (int i = 0; i < nw; i++) d += Minv[i];
		*sm_slot = d;
	}
	for (int j = tid; j < nw; j += nth) { li[j * nw + j] = fv[j]; ui[j * nw + j] = 1.0; }
	__syncthreads();
	*det_out = *sm_slot;
	/* x_j: += -(l_ji f_j) x_i: for i < j, columns 0..i (matrix.cc:573-577); the second factor is U^T, f = 1 */
	for (int i = 0; i + 1 < nw; i++) {
		for (int j = i + 1 + row0; j < nw; j += rstep) {
			const double fl = -(Wk[j * nw + i] * fv[j]), fu = -(Wk[i * nw + j] * 1.0);
			for (int k = col0; k <= i; k += W) {
				li[j * nw + k] = fma(fl, li[i * nw + k], li[j * nw + k]);
				ui[j * nw + k] = fma(fu, ui[i * nw + k], ui[j * nw + k]);
			}
		}
		__syncthreads();
	}
	/* Minv[j][i] = sum_{k >= j} uinv[k][j] linv[k][i]  (i <= j), mirrored: matrix.cc:55-61 */
	for (int j = row0; j < nw; j += rstep)
		for (int i = col0; i <= j; i += W) {
			double s = 0;
			for (int k = j; k < nw; k++) s = fma(ui[k * nw + j], li[k * nw + i], s);
			Minv[j * nw + i] = s; Minv[i * nw + j] = s;
		}
	__syncthreads();
}

/* state_mvn_approx.cc:1133-1155 valid_matrix */
__device__ void ap_valid(double *M, const double *mean, double *sigma, const int n, const int logW)
{
	const int tid = threadIdx.x, nth = blockDim.x;
	const double diag_min = 0.001, cormax = 0.8;
	for (int c = tid; c < n; c += nth) {
		double v = M[c * n + c];
		const double mm = mean[c] * mean[c];
		if (v > mm) v = mm;
		if (v < diag_min) v = diag_min;
		M[c * n + c] = v;
		sigma[c] = sqrt(v);
	}
	__syncthreads();
	const int W = 1 << logW, col0 = tid & (W - 1);
	for (int c = tid >> logW; c < n; c += nth >> logW) {
		const double sc = cormax * sigma[c];
		for (int cc = col0; cc < n; cc += W)
			if (cc > c) {
				const double sigmult = sc * sigma[cc], val = M[c * n + cc];
				if (val > sigmult) { M[c * n + cc] = sigmult; M[cc * n + c] = sigmult; }
				if (val < -sigmult) { M[c * n + cc] = -sigmult; M[cc * n + c] = -sigmult; }
			}
	}
	__syncthreads();
}

/* the rows and columns of the susceptible compartments from the block of the others: state_mvn_approx.cc:1056-1073 */
__device__ void ap_fill_S(const DevApprox &x, double *M)
{
	const int tid = threadIdx.x, nth = blockDim.x, n = x.n;
	for (int e = tid; e < x.nws * n; e += nth) {
		const int j = e / n, q = e - j * n, cj = x.ws[j];
		if (q < x.nws) {
			const int ci = x.ws[q];
			double val = 0.0;
			for (int a = x.wsl_off[j]; a < x.wsl_off[j + 1]; a++)
				for (int b = x.wsl_off[q]; b < x.wsl_off[q + 1]; b++) val += M[x.wsl[a] * n + x.wsl[b]];
			M[cj * n + ci] = val;
		} else {
			const int ci = x.wo[q - x.nws];
			double val = 0.0;
			for (int a = x.wsl_off[j]; a < x.wsl_off[j + 1]; a++) val += M[x.wsl[a] * n + ci];
			M[cj * n + ci] = -val; M[ci * n + cj] = -val;
		}
	}
	__syncthreads();
}

/* state_mvn_approx.cc:688-817.  Returns dL (the same value in every thread); NaN is reported as -LARGE. */
__device__ double ap_observation(const DevModel &m, const DevBatch &b, const DevApprox &x, const int p, const int slice, const double invT,
                                 double *popbase, double *covar, double *lu, double *vecs, const int vstride, double *sm)
{
	const int tid = threadIdx.x, nth = blockDim.x;
	const int n = x.n, nw = x.nwo, sett = x.sl_sett[slice], T = m.T;
	const double *ds = b.disc_spline + (size_t)p * m.nspline * T;
	double *pv = popbase + (size_t)sett * n;
	double *M1 = lu, *l = lu + (size_t)nw * nw, *u = lu + 2 * (size_t)nw * nw, *li = lu + 3 * (size_t)nw * nw, *ui = lu + 4 * (size_t)nw * nw, *Mi = lu + 5 * (size_t)nw * nw;
	double *mu1 = vecs + AV_MU1 * vstride, *Minv_mu = vecs + AV_MINVMU * vstride, *mu = vecs + AV_MUN * vstride, *vec = vecs + AV_VEC * vstride, *tmp = vecs + AV_TMP * vstride;
	for (int i = tid; i < nw; i += nth) { mu1[i] = pv[x.wo[i]]; Minv_mu[i] = 0; }
	for (int e = tid; e < nw * nw; e += nth) { const int j = e / nw, i = e - j * nw; M1[e] = covar[x.wo[j] * n + x.wo[i]]; }
	__syncthreads();
	double det1, det;
	ap_invert_det(M1, Mi, l, u, li, ui, nw, x.logW, &det1, sm);          /* Mi = Minv */
	double mu2 = 0.0, log_factor = 0.0;                          /* kept by thread 0 */
	for (int k = x.sl_off[slice]; k < x.sl_off[slice + 1]; k++) {
		const int ob = x.sl_obs[k];
		double factor = m.obs_factor[ob];
		const int spl = m.obs_fs[ob];
		if (spl >= 0) { int tt = sett; if (tt == T) tt--; factor *= ds[spl * T + tt]; }
		double value = m.obs_value[ob] / factor;
		if (x.obs_type[ob] == 0) {                                 /* transitions: the population at the start of the datum is added */
			const double *p0 = popbase + (size_t)x.obs_sett_i[ob] * n;
			for (int q = x.ocl_off[ob]; q < x.ocl_off[ob + 1]; q++) value += p0[x.ocl[q]];
		}
		const int *cl = x.ocw + x.ocw_off[ob];
		const int ncl = x.ocw_off[ob + 1] - x.ocw_off[ob];
		const double var = x.var_approx[ob] / (invT * factor * factor);
		const double val = 1.0 / var;
		const double value_scale = value / ncl;
		for (int e = tid; e < ncl * ncl; e += nth) { const int a = e / ncl, c2 = e - a * ncl; Mi[cl[a] * nw + cl[c2]] += val; }
		for (int a = tid; a < ncl; a += nth) {
			const int cc = cl[a];
			double acc = Minv_mu[cc], t2 = 0.0;
			const double xa = value_scale - mu1[cc];
			for (int c2 = 0; c2 < ncl; c2++) { const double y = value_scale - mu1[cl[c2]]; acc += val * y; t2 += xa * val * y; }
			Minv_mu[cc] = acc;
			tmp[a] = t2;
		}
		__syncthreads();
		if (tid == 0) {
			for (int a = 0; a < ncl; a++) mu2 += tmp[a];
			log_factor += log(val) - log(2 * 3.14159265358979323846);
		}
		__syncthreads();
	}
	ap_invert_det(Mi, M1, l, u, li, ui, nw, x.logW, &det, sm);           /* M1 = the combined covariance M */
	for (int j = tid; j < nw; j += nth) { double s = 0; for (int i = 0; i < nw; i++) s += M1[j * nw + i] * Minv_mu[i]; mu[j] = s; }
	__syncthreads();
	for (int i = tid; i < nw; i += nth) { double v = mu[i] + mu1[i]; if (v < 0) v = 0; vec[x.wo[i]] = v; }
	for (int e = tid; e < n * n; e += nth) covar[e] = 0;
	__syncthreads();
	for (int i = tid; i < x.nws; i += nth) {
		double sum = 0.0;
		for (int a = x.wsl_off[i]; a < x.wsl_off[i + 1]; a++) sum += vec[x.wsl[a]];
		const int ca = x.ws[i];
		vec[ca] = x.total_pop[(ca / (m.C * m.D)) * m.D + ca % m.D] - sum;
	}
	for (int e = tid; e < nw * nw; e += nth) { const int j = e / nw, i = e - j * nw; covar[x.wo[j] * n + x.wo[i]] = M1[e]; }
	__syncthreads();
	ap_fill_S(x, covar);
	ap_valid(covar, vec, vecs + AV_SIG * vstride, n, x.logW);
	for (int i = tid; i < n; i += nth) pv[i] = vec[i];
	double dL = 0;
	if (tid == 0) {
		double mu_Minv_mu = 0;
		for (int i = 0; i < nw; i++) mu_Minv_mu += mu[i] * Minv_mu[i];
		dL = 0.5 * (log_factor - det1 - det) - 0.5 * (0 + mu2 - mu_Minv_mu);
		if (isnan(dL)) dL = -AP_LARGE;
		else if (dL > 0) dL = 0;
	}
	return ap_bcast(dL, sm);
}

__global__ void __launch_bounds__(AP_THREADS) k_likelihood_approx(const DevModel m, const DevBatch b, const DevApprox x, const double invT)
{
	__shared__ double nts[32 * 32];
	__shared__ double sm_slot[2];
	const int p = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
	const int n = x.n, nt = x.nt, T = m.T, A = m.A, C = m.C, J = m.J, D = m.D, G = m.G;
	const int vstride = n > nt ? n : nt;
	extern __shared__ __align__(16) double ap_dyn[];
	double *popbase = x.pop + (size_t)p * (T + 1) * n;
	double *vecs = x.sm_vec ? ap_dyn : x.vec + (size_t)p * AP_NVEC * vstride;
	double *covar = x.sm_mat ? ap_dyn + AP_NVEC * vstride : x.covar + (size_t)p * n * n;
	double *R = x.sm_mat ? covar + n * n : x.R + (size_t)p * nt * n;
	double *lu = x.lu + (size_t)p * 6 * x.nwo * x.nwo;
	double *Sm = lu;                                              /* n x n, free between observations (6 nwo^2 >= n^2) */
	double *jval = x.jval + (size_t)p * x.jcap;
	/* (row, column) passes: a row per group of W = 2^logW lanes (W = 32 from 32 columns on), no index divisions */
	const int logW = x.logW, W = 1 << logW, col0 = tid & (W - 1);
	const int row0 = tid >> logW, rstep = nth >> logW;
	int *jcol = x.jcol + (size_t)p * x.jcap;
	double *dinf = vecs + AV_DINF * vstride, *Iv = vecs + AV_IV * vstride, *nmi = vecs + AV_NMI * vstride, *bes = vecs + AV_BE * vstride,
	       *mu = vecs + AV_MU * vstride, *dS = vecs + AV_DS * vstride;
	const double *ds = b.disc_spline + (size_t)p * m.nspline * T;
	const double *sus = b.sus + (size_t)p * D, *transrate = b.transrate + (size_t)p * J * D, *infco = b.infco + (size_t)p * C;
	const int *il = b.inflist + (size_t)p * (C + 1);
	const int ninf = il[0];
	const int itr = m.itr, sfrom = m.trans_from[itr];
	const double dt = m.dt;
	bool failed = false;
	for (int k = 1; k <= ninf; k++) if (il[k] == sfrom) failed = true;     /* "The start compartment cannot be infectious", state_mvn_approx.cc:1238 */

	for (int i = tid; i < n; i += nth) popbase[i] = m.pop_init[i];
	for (int e = tid; e < n * n; e += nth) covar[e] = 0;
	for (int e = tid; e < G * G; e += nth) nts[e] = m.N0[e];
	__syncthreads();

	double L = 0.0;
	int o = 0;
	for (int sett = 0; sett <= T && !failed; sett++) {
		if (o < x.nslice && sett == x.sl_sett[o]) {
			if (sett > 0) {
				const double dL = ap_observation(m, b, x, p, o, invT, popbase, covar, lu, vecs, vstride, sm_slot);
				if (dL == -AP_LARGE) { failed = true; break; }
				L += dL;
			}
			o++;
		}
		if (sett == T) break;
		const double *pop = popbase + (size_t)sett * n;
		double *popn = popbase + (size_t)(sett + 1) * n;

		/* ---- transition means of the step (state.cc:223-290, 547-578, 380-456) ---- */
		for (int q = tid; q < A * G; q += nth) {
			const int c = q / G, a = q - c * G;
			double sum = 0;
			for (int k = 1; k <= ninf; k++) {
				const int co = il[k];
				const double inf = infco[co];
				for (int dp = 0; dp < D; dp++)
					if (m.dp_age[dp] == a) { const double pp = pop[(c * C + co) * D + dp]; if (pp != 0) sum += inf * pp; }
			}
			dinf[q] = sum;
		}
		if (m.nmm != 0) for (int e = tid; e < G * G; e += nth) { const int a = e / G, aa = e - a * G; nts[e] = ntime_elem(m, ds, sett, a, aa, m.N0[e]); }
		__syncthreads();
		{
			const double d = ds[m.geo_spline * T + sett], omd = 1 - d;
			for (int q = tid; q < A * G; q += nth) {
				const int c = q / G, a = q - c * G;
				const AreaInfo ai = m.area_info[c];
				double Imap = 0;
				for (int j = 0; j < ai.nn; j++) { const Nbr nb = m.nbr[ai.j0 + j]; Imap += dinf[(nb.off >> 3) + a] * nb.val; }
				const double Idiag = dinf[q] * ai.diag;
				if (G == 1) { double v = Idiag * (d + ai.fac * omd) + Imap * d; if (v < 0) v = 0; nmi[q] = v; }
				else Iv[q] = (Idiag + Imap) * d + ai.fac * Idiag * omd;
				if (a == 0) {
					double be = b.betaR[((size_t)p * m.nR + ai.rg) * T + sett];
					be *= b.areafactor[((size_t)p * m.period + sett / m.spu) * A + c];
					bes[c] = be;
				}
			}
		}
		__syncthreads();
		if (G > 1) {
			for (int q = tid; q < A * G; q += nth) {
				const int c = q / G, a = q - c * G;
				double sum = 0;
				for (int aa = 0; aa < G; aa++) sum += nts[a * G + aa] * Iv[c * G + aa];
				if (sum < 0) sum = 0;
				nmi[q] = sum;
			}
			__syncthreads();
		}
		for (int ta = tid; ta < nt; ta += nth) {
			const int c = ta / (J * D), tr = (ta / D) % J, dp = ta - (c * J + tr) * D;
			const double popu = pop[(c * C + m.trans_from[tr]) * D + dp];
			double mean;
			if (tr == itr) {
				const int a = m.dp_age[dp];
				const AreaInfo ai = m.area_info[c];
				const double et = b.etab[((size_t)p * m.nefoi + ai.ie0) * T + sett];
				const double foi = bes[c] * nmi[c * G + a] + et * m.efoi_agedist[ai.ie0 * G + a];
				dS[c * D + dp] = dt * sus[dp] * foi;                 /* d mean / d S, state_mvn_approx.cc:1225-1231 */
				mean = (popu <= 0) ? 0 : dt * popu * sus[dp] * foi;
			} else mean = (popu <= 0) ? 0 : dt * popu * transrate[tr * D + dp];
			mu[ta] = mean;
		}
		__syncthreads();

		/* ---- the infection rows of the Jacobian (state_mvn_approx.cc:1213-1266): one thread per row lists its
		 * non-zeros (column, value) in ascending compartment order = the reference's order of sums ---- */
		for (int r = tid; r < A * D; r += nth) {
			const int c2 = r / D, dp2 = r - c2 * D, a2 = m.dp_age[dp2];
			const AreaInfo ai = m.area_info[c2];
			const int len = (ai.nn + 1) * ninf * D + 1;
			int k = D * ((ai.j0 + c2) * ninf * D + c2) + dp2 * len;
			const double pref = dt * sus[dp2] * bes[c2];
			const double pS = pop[(c2 * C + sfrom) * D + dp2];
			const int colS = (c2 * C + sfrom) * D + dp2;
			const double valS = dS[c2 * D + dp2];
			int j = 0;
			bool self_done = false;
			while (true) {
				int c; double va;
				const int cn = (j < ai.nn) ? (int)(m.nbr[ai.j0 + j].off >> 3) / G : A;
				if (!self_done && c2 < cn) { c = c2; va = ai.diag; self_done = true; }
				else if (j < ai.nn) { c = cn; va = m.nbr[ai.j0 + j].val; j++; }
				else break;
				bool s_done = (c != c2);
				for (int q = 1; q <= ninf; q++) {
					const int co = il[q];
					const double inf = infco[co];
					if (!s_done && co > sfrom) { jcol[k] = colS; jval[k] = valS; k++; s_done = true; }
					for (int dp = 0; dp < D; dp++) {
						const double valu = pref * nts[m.dp_age[dp] * G + a2] * va * inf;
						jcol[k] = (c * C + co) * D + dp; jval[k] = valu * pS; k++;
					}
				}
				if (!s_done) { jcol[k] = colS; jval[k] = valS; k++; }
			}
		}
		__syncthreads();
		/* ---- R = J covar (covar_matrix_mult_SIMD, matrix.cc:211-236): a row of R per group of W lanes ---- */
		for (int ta = row0; ta < nt; ta += rstep) {
			const int c2 = ta / (J * D), tr = (ta / D) % J, dp2 = ta - (c2 * J + tr) * D;
			double *Rrow = R + ta * n;
			if (tr != itr) {
				const double val = dt * transrate[tr * D + dp2];
				const double *crow = covar + ((c2 * C + m.trans_from[tr]) * D + dp2) * n;
				for (int i = col0; i < n; i += W) Rrow[i] = (val != 0) ? fma(val, crow[i], 0.0) : 0.0;
			} else {
				const AreaInfo ai = m.area_info[c2];
				const int len = (ai.nn + 1) * ninf * D + 1;
				const int k0 = D * ((ai.j0 + c2) * ninf * D + c2) + dp2 * len;
				const double *jv = jval + k0;
				const int *jc = jcol + k0;
				for (int i = col0; i < n; i += 4 * W) {
					/* four columns per lane: the row's (column, value) entries are fetched once for all of them */
					const int i1 = i + W, i2 = i + 2 * W, i3 = i + 3 * W;
					const bool h1 = i1 < n, h2 = i2 < n, h3 = i3 < n;
					double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 2
					for (int k = 0; k < len; k++) {
						const double val = jv[k];
						if (val != 0) {
							const double *crow = covar + jc[k] * n;
							r0 = fma(val, crow[i], r0);
							if (h1) r1 = fma(val, crow[i1], r1);
							if (h2) r2 = fma(val, crow[i2], r2);
							if (h3) r3 = fma(val, crow[i3], r3);
						}
					}
					Rrow[i] = r0;
					if (h1) Rrow[i1] = r1;
					if (h2) Rrow[i2] = r2;
					if (h3) Rrow[i3] = r3;
				}
			}
		}
		__syncthreads();

		/* ---- S = Tm R: row j is the signed sum of the rows of R of j's transitions (covar_matrix_mult_SIMD with the
		 * incidence matrix, state_mvn_approx.cc:1098) ---- */
		for (int j = row0; j < n; j += rstep) {
			const int q0 = x.inc_off[j], q1 = x.inc_off[j + 1];
			for (int i = col0; i < n; i += W) {
				double sj = 0;
				for (int q = q0; q < q1; q++) { const int ta = x.inc_ta[q]; sj += (x.ta_to[ta] == j ? 1.0 : -1.0) * R[ta * n + i]; }
				Sm[j * n + i] = sj;
			}
		}
		__syncthreads();
		/* ---- covar += mu terms (state_mvn_approx.cc:1083-1094) + S + S^T (covar_add_mat, matrix.cc:199-208) ---- */
		for (int j = row0; j < n; j += rstep) {
			const int q0 = x.inc_off[j], q1 = x.inc_off[j + 1];
			for (int i = col0; i < n; i += W) {
				double v = covar[j * n + i];
				for (int q = q0; q < q1; q++) {
					const int ta = x.inc_ta[q];
					const int fr = x.ta_from[ta], to = x.ta_to[ta];
					if (i == j) v += mu[ta];
					else if ((fr == j ? to : fr) == i) v -= mu[ta];
				}
				v += Sm[j * n + i] + Sm[i * n + j];
				covar[j * n + i] = v;
			}
		}
		__syncthreads();
		ap_valid(covar, pop, vecs + AV_SIG * vstride, n, x.logW);

		/* ---- state.cc:460-485 update_pop with transnum = transmean ---- */
		for (int ca = tid; ca < n; ca += nth) {
			double v = pop[ca];
			for (int q = x.inc_off[ca]; q < x.inc_off[ca + 1]; q++) {
				const int ta = x.inc_ta[q];
				const double num = mu[ta];
				if (num != 0) { if (x.ta_from[ta] == ca) v -= num; else v += num; }
			}
			popn[ca] = v;
		}
		__syncthreads();
	}
	if (tid == 0) x.L[p] = failed ? -AP_LARGE : L;
}

#endif
