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
 * Mapping: ONE CTA PER PARTICLE, the whole time loop inside the kernel (no launch per step); the particle's
 * covariance, the product R = J covar and the LU factors live in a per-particle scratch in HBM that the CTA
 * re-reads through L1/L2 (n <= a few hundred: tens to hundreds of KB); every phase is a block-wide pass over
 * n x n (or transitions x n) outputs with unit-stride lanes, separated by __syncthreads.  The Jacobian is never
 * formed: its few non-zeros per row are recomputed from the step's tables while R is accumulated (ascending
 * compartment index = the reference's order of sums).  Sums that the reference does sequentially are sequential
 * here too (one thread per output), so results agree with the CPU restatement to rounding.
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

/* matrix.cc:18-62 + 564-579: Minv = M^-1 by Crout LU without pivoting (a zero pivot becomes TINY), det = sum log|l_ii|.
 * M, Minv, l, u, li, ui are nw x nw in the particle's scratch; every thread of the CTA calls this. */
__device__ void ap_invert_det(const double *M, double *Minv, double *l, double *u, double *li, double *ui, const int nw, double *det_out, double *sm_slot)
{
	const int tid = threadIdx.x, nth = blockDim.x;
	for (int i = 0; i < nw; i++) {
		for (int j = i + tid; j < nw; j += nth) {
			double s = 0;
			for (int k = 0; k < i; k++) s += l[j * nw + k] * u[i * nw + k];
			l[j * nw + i] = M[j * nw + i] - s;
		}
		__syncthreads();
		double val = l[i * nw + i];
		if (val == 0) val = AP_TINY;
		if (tid == 0) u[i * nw + i] = 1;
		for (int j = i + 1 + tid; j < nw; j += nth) {
			double s = 0;
			for (int k = 0; k < i; k++) s += l[i * nw + k] * u[j * nw + k];
			u[j * nw + i] = (M[i * nw + j] - s) / val;
		}
		__syncthreads();
	}
	if (tid == 0) {
		double d = 0;
		for (int i = 0; i < nw; i++) { double vall = l[i * nw + i]; if (vall < 0) vall = -vall; d += log(1.0 * vall); }
		*sm_slot = d;
	}
	/* inverses of the two lower-triangular factors, one column per thread (the reference's row recurrence,
	 * matrix.cc:564-579, element by element: x_jk = [j==k] f_j - sum_{i<j} (l_ji f_j) x_ik, f_j = 1/l_jj) */
	for (int k = tid; k < 2 * nw; k += nth) {
		const bool second = k >= nw;
		const int kk = second ? k - nw : k;
		const double *Lm = second ? u : l;
		double *X = second ? ui : li;
		for (int j = 0; j < kk; j++) X[j * nw + kk] = 0;
		for (int j = kk; j < nw; j++) {
			const double f = 1.0 / Lm[j * nw + j];
			double acc = (j == kk) ? f : 0.0;
			for (int i = kk; i < j; i++) acc += (-(Lm[j * nw + i] * f)) * X[i * nw + kk];
			X[j * nw + kk] = acc;
		}
	}
	__syncthreads();
	*det_out = *sm_slot;
	/* Minv[j][i] = sum_{k >= j} uinv[k][j] linv[k][i]  (i <= j), mirrored: matrix.cc:55-61 */
	for (int e = tid; e < nw * nw; e += nth) {
		const int j = e / nw, i = e - j * nw;
		if (i <= j) {
			double s = 0;
			for (int k = j; k < nw; k++) s += ui[k * nw + j] * li[k * nw + i];
			Minv[j * nw + i] = s; Minv[i * nw + j] = s;
		}
	}
	__syncthreads();
}

/* state_mvn_approx.cc:1133-1155 valid_matrix */
__device__ void ap_valid(double *M, const double *mean, double *sigma, const int n)
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
	for (int e = tid; e < n * n; e += nth) {
		const int c = e / n, cc = e - c * n;
		if (cc > c) {
			const double sigmult = cormax * sigma[c] * sigma[cc], val = M[c * n + cc];
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
	ap_invert_det(M1, Mi, l, u, li, ui, nw, &det1, sm);          /* Mi = Minv */
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
	ap_invert_det(Mi, M1, l, u, li, ui, nw, &det, sm);           /* M1 = the combined covariance M */
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
	ap_valid(covar, vec, vecs + AV_SIG * vstride, n);
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
	double *popbase = x.pop + (size_t)p * (T + 1) * n;
	double *covar = x.covar + (size_t)p * n * n, *R = x.R + (size_t)p * nt * n;
	double *lu = x.lu + (size_t)p * 6 * x.nwo * x.nwo;
	double *vecs = x.vec + (size_t)p * AP_NVEC * vstride;
	double *dinf = vecs + AV_DINF * vstride, *Iv = vecs + AV_IV * vstride, *nmi = vecs + AV_NMI * vstride, *bes = vecs + AV_BE * vstride,
	       *mu = vecs + AV_MU * vstride, *dS = vecs + AV_DS * vstride;
	const double *ds = b.disc_spline + (size_t)p * m.nspline * T;
	const double *sus = b.sus + (size_t)p * D, *transrate = b.transrate + (size_t)p * J * D, *infco = b.infco + (size_t)p * C;
	const int *il = b.inflist + (size_t)p * (C + 1);
	const int ninf = il[0];
	const int itr = m.itr, sfrom = m.trans_from[itr];
	const double dt = m.dt;

	for (int i = tid; i < n; i += nth) popbase[i] = m.pop_init[i];
	for (int e = tid; e < n * n; e += nth) covar[e] = 0;
	for (int e = tid; e < G * G; e += nth) nts[e] = m.N0[e];
	__syncthreads();

	double L = 0.0;
	int o = 0;
	bool failed = false;
	for (int sett = 0; sett <= T; sett++) {
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

		/* ---- R = J covar, the Jacobian's non-zeros recomputed on the fly in ascending compartment order ---- */
		for (int e = tid; e < nt * n; e += nth) {
			const int ta = e / n, i = e - ta * n;
			const int c2 = ta / (J * D), tr = (ta / D) % J, dp2 = ta - (c2 * J + tr) * D;
			double r = 0;
			if (tr != itr) {
				const double val = dt * transrate[tr * D + dp2];
				if (val != 0) r += val * covar[((c2 * C + m.trans_from[tr]) * D + dp2) * n + i];
			} else {
				const int a2 = m.dp_age[dp2];
				const AreaInfo ai = m.area_info[c2];
				const double pref = dt * sus[dp2] * bes[c2];
				const double pS = pop[(c2 * C + sfrom) * D + dp2];
				int j = 0;
				bool self_done = false;
				while (true) {
					int c; double va;
					const int cn = (j < ai.nn) ? (int)(m.nbr[ai.j0 + j].off >> 3) / G : A;
					if (!self_done && c2 < cn) { c = c2; va = ai.diag; self_done = true; }
					else if (j < ai.nn) { c = cn; va = m.nbr[ai.j0 + j].val; j++; }
					else break;
					for (int co = 0; co < C; co++) {
						if (co == sfrom) {
							if (c == c2) { const double val = dS[c2 * D + dp2]; if (val != 0) r += val * covar[((c2 * C + sfrom) * D + dp2) * n + i]; }
							continue;
						}
						const double inf = infco[co];
						if (inf == 0) continue;
						for (int dp = 0; dp < D; dp++) {
							const double valu = pref * nts[m.dp_age[dp] * G + a2] * va * inf;
							const double val = valu * pS;
							if (val != 0) r += val * covar[((c * C + co) * D + dp) * n + i];
						}
					}
				}
			}
			R[e] = r;
		}
		__syncthreads();

		/* ---- covar += mu terms + S + S^T, S[j][:] = sum over the transitions of j of +-R[ta][:]  (state_mvn_approx.cc:1083-1101) ---- */
		for (int e = tid; e < n * n; e += nth) {
			const int j = e / n, i = e - j * n;
			double v = covar[e];
			double sji = 0, sij = 0;
			for (int q = x.inc_off[j]; q < x.inc_off[j + 1]; q++) {
				const int ta = x.inc_ta[q];
				const int fr = x.ta_from[ta], to = x.ta_to[ta];
				const double val = mu[ta];
				if (i == j) v += val;
				else if ((fr == j ? to : fr) == i) v -= val;
				sji += (to == j ? 1.0 : -1.0) * R[ta * n + i];
			}
			for (int q = x.inc_off[i]; q < x.inc_off[i + 1]; q++) {
				const int ta = x.inc_ta[q];
				sij += (x.ta_to[ta] == i ? 1.0 : -1.0) * R[ta * n + j];
			}
			v += sji + sij;
			covar[e] = v;
		}
		__syncthreads();
		ap_valid(covar, pop, vecs + AV_SIG * vstride, n);

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
