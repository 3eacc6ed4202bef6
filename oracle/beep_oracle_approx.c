/* beep_oracle_approx.c -- TEST INFRASTRUCTURE (included by beep_oracle.c; never on the product path).
 *
 * CPU restatement of the reference's MVN-approximate likelihood, the objective of mode=map
 * (ML::generate_samples, ml.cc:389-413):
 *
 *     State::likelihood_approx        state_mvn_approx.cc:11-91
 *     State::apply_observation_fast   state_mvn_approx.cc:688-817
 *     State::covar_change_fast        state_mvn_approx.cc:1081-1111
 *     State::set_transmean_gradients  state_mvn_approx.cc:1199-1284
 *     State::valid_matrix, correct_mu_covar, remove_S / add_S, set_mu / set_p / put_p
 *                                     state_mvn_approx.cc:946-1076, 1133-1195
 *     State::initialise_approx        state_mvn_approx.cc:1288-1475  (index lists)
 *     ObservationModel::generate_obs_slice  obsmodel.cc:357-385
 *     Data::normal_approximation      data.cc:1085-1115
 *     Model::trans_comp_divide        model.cc:3524-3566
 *     invert_determinant_SIMD, invert_lower_triangular, covar_matrix_mult_SIMD   matrix.cc:18-62, 564-579, 211-236
 *
 * Plain sequential loops; sums run in index order, with the fused multiply-adds the reference's AVX routines use
 * written as fma() (the reference's dot products keep four partial sums, so agreement with the reference is to
 * rounding, not bit for bit: tests/test_approx_*.py state the tolerance).  Two deliberate differences, both equal to the reference in exact arithmetic:
 *   - the infectivity map is recomputed from the mean population at every step (the reference updates it from
 *     transnum[sett-1], to which apply_observation_fast adds the transitions Tinv_without_S (vec - mu_start) that
 *     reproduce the population change of the observation, state_mvn_approx.cc:803-814: the two maps are equal);
 *     those corrected transnum are therefore not formed here -- they do not enter the likelihood otherwise;
 *   - the gradient matrix is cleared at the start of a call (the reference keeps entries of earlier calls whose
 *     infectivity parameter has since become zero; no model of the reference varies an infectivity).
 * As in the reference the method needs one strain, as many transitions as non-susceptible compartments, and
 * observations that do not measure the susceptible compartment; otherwise -1e300 is returned.
 */

#define AP_BAD (-1e300)

typedef struct ApproxIdx {
	int n, nt, nwo, nws, start;
	int *wo, *ws, *wsl_off, *wsl;     /* comp_without_S, comp_with_S, comp_with_S_list (CSR, ca numbering) */
	double *total_pop;                /* [A][D] */
	int *ocl_off, *ocl;               /* obs_clist (ca)                          */
	int *ocw_off, *ocw;               /* obs_clist_without_S (index into wo)     */
	double *var_approx;               /* [nobs] data.cc:1085-1115                */
	int nslice, *sl_sett, *sl_off, *sl_obs;
} ApproxIdx;

#define CA(c, co, dp) ((((c) * m->C) + (co)) * m->D + (dp))
#define TA(c, tr, dp) ((((c) * m->J) + (tr)) * m->D + (dp))

/* model.cc:3524-3566: the compartments on the far side (dir = 1) of transition tr */
static int trans_comp_divide(const BeepModelDesc *m, int tr, int *clist)
{
	int map[64];
	for (int c = 0; c < m->C; c++) map[c] = -1;
	map[m->trans_from[tr]] = 0; map[m->trans_to[tr]] = 1;
	int flag;
	do {
		flag = 0;
		for (int k = 0; k < m->J; k++) {
			int from = map[m->trans_from[k]], to = map[m->trans_to[k]];
			int num = (from < 0) + (to < 0);
			if (num == 0) {
				if (from != to && !(m->trans_from[k] == m->trans_from[tr] && m->trans_to[k] == m->trans_to[tr])) return -1;
			} else if (num == 1) {
				if (from < 0) map[m->trans_from[k]] = to; else map[m->trans_to[k]] = from;
				flag = 1;
			}
		}
	} while (flag);
	int n = 0;
	for (int c = 0; c < m->C; c++) if (map[c] == 1) clist[n++] = c;
	return n;
}

static void approx_free(ApproxIdx *x)
{
	free(x->wo); free(x->ws); free(x->wsl_off); free(x->wsl); free(x->total_pop); free(x->ocl_off); free(x->ocl);
	free(x->ocw_off); free(x->ocw); free(x->var_approx); free(x->sl_sett); free(x->sl_off); free(x->sl_obs);
}

/* state_mvn_approx.cc:1288-1475 initialise_approx + obsmodel.cc:357-385 generate_obs_slice */
static int approx_init(const BeepModelDesc *m, ApproxIdx *x)
{
	memset(x, 0, sizeof *x);
	if (m->S != 1 || m->C > 64 || m->n_table_type != m->ntable) return -1;
	const int A = m->A, C = m->C, D = m->D, J = m->J;
	x->n = A * C * D; x->nt = A * J * D;
	x->start = m->trans_from[m->infection_trans];                         /* model.cc:197 */
	x->wo = (int *)calloc(x->n, sizeof(int)); x->ws = (int *)calloc(x->n, sizeof(int));
	int *wo_of = (int *)calloc(x->n, sizeof(int));
	for (int ca = 0; ca < x->n; ca++) {
		int co = (ca / D) % C;
		if (co != x->start) { wo_of[ca] = x->nwo; x->wo[x->nwo++] = ca; } else { wo_of[ca] = -1; x->ws[x->nws++] = ca; }
	}
	if (x->nwo != x->nt) { free(wo_of); return -1; }                      /* state_mvn_approx.cc:1376 */
	x->wsl_off = (int *)calloc(x->nws + 1, sizeof(int)); x->wsl = (int *)calloc(x->nwo + 1, sizeof(int));
	int cnt = 0;
	for (int i = 0; i < x->nws; i++) {
		int cai = x->ws[i], ci = cai / (C * D), dpi = cai % D;
		x->wsl_off[i] = cnt;
		for (int j = 0; j < x->nwo; j++) { int caj = x->wo[j]; if (caj / (C * D) == ci && caj % D == dpi) x->wsl[cnt++] = caj; }
	}
	x->wsl_off[x->nws] = cnt;
	x->total_pop = dalloc((long)A * D);
	for (int c = 0; c < A; c++)
		for (int dp = 0; dp < D; dp++) {
			double sum = 0.0;
			for (int co = 0; co < C; co++) sum += m->pop_init[((long)c * C + co) * D + dp];
			x->total_pop[c * D + dp] = sum;
		}
	/* observation compartment lists */
	x->ocl_off = (int *)calloc(m->nobs + 1, sizeof(int)); x->ocw_off = (int *)calloc(m->nobs + 1, sizeof(int));
	long cap = 0;
	for (int o = 0; o < m->nobs; o++) cap += (long)C * (m->obs_area_off[o + 1] - m->obs_area_off[o]) * (m->obs_dp_off[o + 1] - m->obs_dp_off[o]);
	x->ocl = (int *)calloc(cap + 1, sizeof(int)); x->ocw = (int *)calloc(cap + 1, sizeof(int));
	x->var_approx = dalloc(m->nobs);
	int n1 = 0, n2 = 0, bad = 0;
	for (int o = 0; o < m->nobs; o++) {
		int tb = m->obs_table[o], type = m->table_type[tb];
		int complist[64], nc = 0;
		if (type == 1 || type == 2) {                                      /* POP, POPFRAC (consts.hh:84) */
			for (int k = m->table_comp_off[tb]; k < m->table_comp_off[tb + 1]; k++) complist[nc++] = m->table_comp[k];
		} else {                                                           /* TRANS, MARGINAL */
			for (int k = m->table_trans_off[tb]; k < m->table_trans_off[tb + 1]; k++) {
				int after[64], na = trans_comp_divide(m, m->table_trans[k], after);
				if (na < 0) bad = 1;
				for (int q = 0; q < na; q++) {
					int dup = 0;
					for (int r = 0; r < nc; r++) if (complist[r] == after[q]) dup = 1;
					if (dup) bad = 1; else complist[nc++] = after[q];
				}
			}
		}
		x->ocl_off[o] = n1; x->ocw_off[o] = n2;
		for (int k = 0; k < nc; k++) {
			if (complist[k] == x->start) bad = 1;
			for (int ia = m->obs_area_off[o]; ia < m->obs_area_off[o + 1]; ia++)
				for (int id = m->obs_dp_off[o]; id < m->obs_dp_off[o + 1]; id++) {
					int ca = CA(m->obs_area[ia], complist[k], m->obs_dp[id]);
					x->ocl[n1++] = ca;
					if (wo_of[ca] >= 0) x->ocw[n2++] = wo_of[ca];
				}
		}
		/* data.cc:1085-1115 */
		double value = m->obs_value[o], va = 0;
		switch (m->obs_model[o]) {
		case 0: case 1: va = m->obs_sd[o] * m->obs_sd[o]; break;           /* NORMAL, NORMAL_PERCENT (consts.hh:40) */
		case 2: { double lam = value; if (lam < 1) lam = 1; va = lam; } break;
		case 3: { double mm = value; if (mm < 1) mm = 1; va = mm + mm * mm / m->obs_shape[o]; } break;
		}
		x->var_approx[o] = va;
	}
	x->ocl_off[m->nobs] = n1; x->ocw_off[m->nobs] = n2;
	free(wo_of);
	if (bad) return -1;
	/* slices: obsmodel.cc:357-385 (slice time = start of a population datum, end of a transition datum) */
	int *osett = (int *)calloc(m->nobs + 1, sizeof(int));
	x->sl_sett = (int *)calloc(m->nobs + 1, sizeof(int)); x->sl_off = (int *)calloc(m->nobs + 2, sizeof(int)); x->sl_obs = (int *)calloc(m->nobs + 1, sizeof(int));
	for (int o = 0; o < m->nobs; o++) { int type = m->table_type[m->obs_table[o]]; osett[o] = (type == 1 || type == 2) ? m->obs_sett_i[o] : m->obs_sett_f[o]; }
	int last = -1, k = 0;
	while (1) {
		int best = -1;
		for (int o = 0; o < m->nobs; o++) if (osett[o] > last && (best < 0 || osett[o] < best)) best = osett[o];
		if (best < 0) break;
		x->sl_sett[x->nslice] = best; x->sl_off[x->nslice] = k;
		for (int o = 0; o < m->nobs; o++) if (osett[o] == best) x->sl_obs[k++] = o;
		x->nslice++; last = best;
	}
	x->sl_off[x->nslice] = k;
	free(osett);
	return 0;
}

/* matrix.cc:564-579 invert_lower_triangular (M is normalised in place); the row update is the reference's fused
 * multiply-add (sub_row_SIMD, matrix.cc:769-795: _mm256_fmadd_pd) */
static void ap_invert_lower(double *M, double *Minv, int n, int ld)
{
	for (int j = 0; j < n; j++) {
		double fac = 1.0 / M[j * ld + j];
		for (int i = 0; i < n; i++) { Minv[j * ld + i] = 0; M[j * ld + i] *= fac; }
		Minv[j * ld + j] = fac;
		M[j * ld + j] = 1;
		for (int i = 0; i < j; i++) {
			double f = -M[j * ld + i];
			for (int k = 0; k <= i; k++) Minv[j * ld + k] = fma(f, Minv[i * ld + k], Minv[j * ld + k]);
		}
	}
}

/* matrix.cc:18-62 invert_determinant_SIMD, DOUBLE branch: Crout LU without pivoting, det = sum log|l_ii|.  The
 * dot products are taken off the matrix element one fused multiply-add at a time in index order (the reference's
 * dot_SIMD keeps four partial sums, matrix.cc:379-410: same algorithm, another order of rounding). */
static void ap_invert_det(const double *M, double *Minv, int n, double *det)
{
	double *l = dalloc((long)n * n), *u = dalloc((long)n * n), *li = dalloc((long)n * n), *ui = dalloc((long)n * n);
	for (int i = 0; i < n; i++) {
		for (int j = i; j < n; j++) {
			double acc = M[j * n + i];
			for (int k = 0; k < i; k++) acc = fma(-l[j * n + k], u[i * n + k], acc);
			l[j * n + i] = acc;
		}
		u[i * n + i] = 1;
		for (int j = i + 1; j < n; j++) {
			double val = l[i * n + i]; if (val == 0) val = TINY;
			double acc = M[i * n + j];
			for (int k = 0; k < i; k++) acc = fma(-l[i * n + k], u[j * n + k], acc);
			u[j * n + i] = acc / val;
		}
	}
	double d = 0;
	for (int i = 0; i < n; i++) {
		double valu = u[i * n + i]; if (valu < 0) valu = -valu;
		double vall = l[i * n + i]; if (vall < 0) vall = -vall;
		d += log(valu * vall);
	}
	*det = d;
	ap_invert_lower(l, li, n, n);
	ap_invert_lower(u, ui, n, n);
	/* Minv[j][i] = sum_{k >= max(i,j)} uinv^T[j][k] linv^T[i][k] = sum_k ui[k][j] li[k][i], mirrored */
	for (int j = 0; j < n; j++)
		for (int i = 0; i <= j; i++) {
			double s = 0;
			for (int k = j; k < n; k++) s = fma(ui[k * n + j], li[k * n + i], s);
			Minv[j * n + i] = s; Minv[i * n + j] = s;
		}
	free(l); free(u); free(li); free(ui);
}

/* state_mvn_approx.cc:1133-1155 valid_matrix */
static void ap_valid_matrix(double *M, const double *mean, int n)
{
	const double diag_min = 0.001, cormax = 0.8;
	double *sigma = dalloc(n);
	for (int c = 0; c < n; c++) {
		if (M[c * n + c] > mean[c] * mean[c]) M[c * n + c] = mean[c] * mean[c];
		if (M[c * n + c] < diag_min) M[c * n + c] = diag_min;
		sigma[c] = sqrt(M[c * n + c]);
	}
	for (int c = 0; c < n; c++) {
		double sig1 = sigma[c];
		for (int cc = c + 1; cc < n; cc++) {
			double sigmult = cormax * sig1 * sigma[cc], val = M[c * n + cc];
			if (val > sigmult) { M[c * n + cc] = sigmult; M[cc * n + c] = sigmult; }
			if (val < -sigmult) { M[c * n + cc] = -sigmult; M[cc * n + c] = -sigmult; }
		}
	}
	free(sigma);
}

/* the with-S rows and columns from the without-S block: state_mvn_approx.cc:1056-1073 (add_S) = :956-975 (correct_mu_covar) */
static void ap_fill_S(const ApproxIdx *x, double *M)
{
	const int n = x->n;
	for (int j = 0; j < x->nws; j++) {
		int cj = x->ws[j];
		for (int i = 0; i < x->nws; i++) {
			int ci = x->ws[i];
			double val = 0.0;
			for (int a = x->wsl_off[j]; a < x->wsl_off[j + 1]; a++)
				for (int b = x->wsl_off[i]; b < x->wsl_off[i + 1]; b++) val += M[x->wsl[a] * n + x->wsl[b]];
			M[cj * n + ci] = val;
		}
		for (int i = 0; i < x->nwo; i++) {
			int ci = x->wo[i];
			double val = 0.0;
			for (int a = x->wsl_off[j]; a < x->wsl_off[j + 1]; a++) val += M[x->wsl[a] * n + ci];
			M[cj * n + ci] = -val; M[ci * n + cj] = -val;
		}
	}
}

/* state_mvn_approx.cc:688-817 apply_observation_fast (pop of step sett and covar are updated in place) */
static double ap_apply_observation(OracleState *s, const ApproxIdx *x, int slice, double *covar, double invT)
{
	const BeepModelDesc *m = s->m;
	const int n = x->n, nwo = x->nwo, sett = x->sl_sett[slice], T = m->T;
	double *pv = &POP(s, sett, 0, 0, 0);                                 /* set_p: ca order IS the [A][C][D] order */
	double *mu1 = dalloc(nwo), *M1 = dalloc((long)nwo * nwo), *Minv = dalloc((long)nwo * nwo), *Minv_mu = dalloc(nwo);
	for (int i = 0; i < nwo; i++) mu1[i] = pv[x->wo[i]];
	for (int j = 0; j < nwo; j++) for (int i = 0; i < nwo; i++) M1[j * nwo + i] = covar[x->wo[j] * n + x->wo[i]];
	double det1 = 0.0, mu_Minv_mu2 = 0.0, log_factor = 0.0;
	ap_invert_det(M1, Minv, nwo, &det1);
	for (int k = x->sl_off[slice]; k < x->sl_off[slice + 1]; k++) {
		const int ob = x->sl_obs[k], type = m->table_type[m->obs_table[ob]];
		double factor = m->obs_factor[ob];
		int spl = m->obs_factor_spline[ob];
		if (spl >= 0) { int tt = sett; if (tt == T) tt--; factor *= s->disc_spline[spl * T + tt]; }
		double value = m->obs_value[ob] / factor;
		if (type == 0) {                                                   /* TRANS: add the population at the start of the datum */
			const double *p0 = &POP(s, m->obs_sett_i[ob], 0, 0, 0);
			for (int q = x->ocl_off[ob]; q < x->ocl_off[ob + 1]; q++) value += p0[x->ocl[q]];
		}
		const int *cl = x->ocw + x->ocw_off[ob];
		const int ncl = x->ocw_off[ob + 1] - x->ocw_off[ob];
		double var = x->var_approx[ob] / (invT * factor * factor);
		double val = 1.0 / var;
		double value_scale = value / ncl;
		for (int a = 0; a < ncl; a++)
			for (int b = 0; b < ncl; b++) {
				int cc = cl[a], cc2 = cl[b];
				Minv[cc * nwo + cc2] += val;
				Minv_mu[cc] += val * (value_scale - mu1[cc2]);
				mu_Minv_mu2 += (value_scale - mu1[cc]) * val * (value_scale - mu1[cc2]);
			}
		log_factor += log(val) - log(2 * M_PI);
	}
	double det = 0.0;
	double *Mn = dalloc((long)nwo * nwo), *mu = dalloc(nwo);
	ap_invert_det(Minv, Mn, nwo, &det);
	for (int j = 0; j < nwo; j++) { double sum = 0; for (int i = 0; i < nwo; i++) sum += Mn[j * nwo + i] * Minv_mu[i]; mu[j] = sum; }   /* matrix.cc:108-126 */
	double mu_Minv_mu = 0; for (int i = 0; i < nwo; i++) mu_Minv_mu += mu[i] * Minv_mu[i];
	/* mu_final = mu + mu1, clamped at 0; add_S (state_mvn_approx.cc:992-1016) */
	double *vec = dalloc(n), *tp = dalloc((long)m->A * m->D);
	memcpy(tp, x->total_pop, (size_t)m->A * m->D * sizeof(double));
	for (int i = 0; i < nwo; i++) {
		double v = mu[i] + mu1[i]; if (v < 0) v = 0;
		int ca = x->wo[i];
		vec[ca] = v;
		tp[(ca / (m->C * m->D)) * m->D + ca % m->D] -= v;
	}
	for (int i = 0; i < x->nws; i++) { int ca = x->ws[i]; vec[ca] = tp[(ca / (m->C * m->D)) * m->D + ca % m->D]; }
	/* covar = add_S(M) */
	for (long q = 0; q < (long)n * n; q++) covar[q] = 0;
	for (int j = 0; j < nwo; j++) for (int i = 0; i < nwo; i++) covar[x->wo[j] * n + x->wo[i]] = Mn[j * nwo + i];
	ap_fill_S(x, covar);
	/* correct_mu_covar (state_mvn_approx.cc:946-976) */
	for (int i = 0; i < x->nws; i++) {
		double sum = 0.0; for (int a = x->wsl_off[i]; a < x->wsl_off[i + 1]; a++) sum += vec[x->wsl[a]];
		int ca = x->ws[i];
		vec[ca] = x->total_pop[(ca / (m->C * m->D)) * m->D + ca % m->D] - sum;
	}
	ap_fill_S(x, covar);
	ap_valid_matrix(covar, vec, n);
	memcpy(pv, vec, (size_t)n * sizeof(double));                         /* put_p */
	double dL = 0.5 * (log_factor - det1 - det) - 0.5 * (0 + mu_Minv_mu2 - mu_Minv_mu);
	free(mu1); free(M1); free(Minv); free(Minv_mu); free(Mn); free(mu); free(vec); free(tp);
	if (isnan(dL)) return -LARGE;
	if (dL > 0) dL = 0;
	return dL;
}

/* state_mvn_approx.cc:1199-1284 set_transmean_gradients: dtm[ta][ca] (dense, nt x n) */
static void ap_gradients(OracleState *s, const ApproxIdx *x, int sett, double *dtm)
{
	const BeepModelDesc *m = s->m;
	const int A = m->A, G = m->G, D = m->D, T = m->T, n = x->n, tr = m->infection_trans, from = m->trans_from[tr];
	const double dt = (double)m->period / T;
	double *NMI = dalloc(G), *eta_age = dalloc(G);
	const double *Nt = s->Ntime + (long)sett * G * G;
	for (int c = 0; c < A; c++) {
		get_NMI(s, sett, 0, c, NMI);
		double be = s->beta[((long)0 * A + c) * T + sett];
		be *= s->areafactor[(sett / m->spu) * A + c];
		int ie = m->efoi_ref[c];
		double et = s->disc_spline[m->efoi_spline_ref[ie] * T + sett] / m->popsize;
		for (int a = 0; a < G; a++) eta_age[a] = et * m->efoi_agedist[ie * G + a];
		for (int dp = 0; dp < D; dp++) {
			int a = m->dp_age[dp];
			dtm[(long)TA(c, tr, dp) * n + CA(c, from, dp)] = dt * s->sus[dp] * (be * NMI[a] + eta_age[a]);
		}
		for (int co = 0; co < m->C; co++) {
			double inf = s->paramval[m->comp_inf_param[co]];
			if (inf != 0) {
				int jmax = m->M_off[c + 1] - m->M_off[c];
				for (int j = 0; j <= jmax; j++) {
					int c2; double va;
					if (j == jmax) { c2 = c; va = m->M_diag[c]; } else { c2 = m->M_to[m->M_off[c] + j]; va = m->M_val[m->M_off[c] + j]; }
					double be2 = s->beta[((long)0 * A + c2) * T + sett];
					be2 *= s->areafactor[(sett / m->spu) * A + c2];
					for (int dp = 0; dp < D; dp++) {
						int a = m->dp_age[dp];
						for (int dp2 = 0; dp2 < D; dp2++) {
							int a2 = m->dp_age[dp2];
							double valu = dt * s->sus[dp2] * be2 * Nt[a * G + a2] * va * inf;
							dtm[(long)TA(c2, tr, dp2) * n + CA(c, co, dp)] = valu * POP(s, sett, c2, from, dp2);
						}
					}
				}
			}
		}
	}
	for (int k = 0; k < m->J; k++)
		if (!m->trans_inf[k])
			for (int c = 0; c < A; c++)
				for (int dp = 0; dp < D; dp++) dtm[(long)TA(c, k, dp) * n + CA(c, m->trans_from[k], dp)] = dt * s->transrate[k * D + dp];
	free(NMI); free(eta_age);
}

/* state_mvn_approx.cc:11-91.  out_pop (may be NULL) receives the mean populations [T+1][A][C][D], out_var the
 * variances (diagonal of the covariance at the start of each step) [T][A][C][D] */
double oracle_likelihood_approx(OracleState *s, const double *theta, double invT, double *out_pop, double *out_var)
{
	const BeepModelDesc *m = s->m;
	ApproxIdx X, *x = &X;
	if (approx_init(m, x) != 0) { approx_free(x); return AP_BAD; }
	const int n = x->n, nt = x->nt, T = m->T, A = m->A, J = m->J, D = m->D;
	oracle_set_param(s, theta);
	pop_init(s);
	double *covar = dalloc((long)n * n), *old = dalloc((long)n * n), *dtm = dalloc((long)nt * n), *R = dalloc((long)nt * n), *P = dalloc((long)n * n);
	double L = 0.0;
	int o = 0;
	for (int sett = 0; sett <= T; sett++) {
		if (o < x->nslice && sett == x->sl_sett[o]) {
			if (sett > 0) {
				double dL = ap_apply_observation(s, x, o, covar, invT);
				if (dL == -LARGE) { L = -LARGE; break; }
				L += dL;
			}
			o++;
		}
		if (sett == T) break;
		memcpy(old, covar, (size_t)n * n * sizeof(double));                /* pop_covar[sett] */
		if (out_var) for (int c = 0; c < n; c++) out_var[(long)sett * n + c] = covar[(long)c * n + c];
		set_I_from_pop(s, sett);
		for (int c = 0; c < A; c++) set_transmean(s, sett, c);
		memcpy(&TN(s, sett, 0, 0, 0), &TM(s, sett, 0, 0, 0), (size_t)nt * sizeof(double));
		ap_gradients(s, x, sett, dtm);
		/* covar_change_fast (state_mvn_approx.cc:1081-1111), forward */
		const double *mu = &TM(s, sett, 0, 0, 0);                          /* set_mu: ta order IS the [A][J][D] order */
		for (int ta = 0; ta < nt; ta++) {
			int c = ta / (J * D), tr = (ta / D) % J, dp = ta % D;
			int cci = CA(c, m->trans_from[tr], dp), ccf = CA(c, m->trans_to[tr], dp);
			double val = mu[ta];
			covar[(long)ccf * n + ccf] += val; covar[(long)cci * n + cci] += val;
			covar[(long)cci * n + ccf] -= val; covar[(long)ccf * n + cci] -= val;
		}
		for (int ta = 0; ta < nt; ta++) {                                  /* R = dtransmean_dp * covar_old (matrix.cc:211-236) */
			double *r = R + (long)ta * n;
			for (int i = 0; i < n; i++) r[i] = 0;
			for (int ca = 0; ca < n; ca++) {
				double val = dtm[(long)ta * n + ca];
				if (val != 0) { const double *row = old + (long)ca * n; for (int i = 0; i < n; i++) r[i] = fma(val, row[i], r[i]); }   /* sub_row_SIMD: fused */
			}
		}
		for (long q = 0; q < (long)n * n; q++) P[q] = 0;                   /* P = T * R, T[ca][ta] = -1 (from) / +1 (to) */
		for (int ta = 0; ta < nt; ta++) {
			int c = ta / (J * D), tr = (ta / D) % J, dp = ta % D;
			int cci = CA(c, m->trans_from[tr], dp), ccf = CA(c, m->trans_to[tr], dp);
			const double *r = R + (long)ta * n;
			double *pi = P + (long)cci * n, *pf = P + (long)ccf * n;
			for (int i = 0; i < n; i++) { pi[i] += -1.0 * r[i]; pf[i] += r[i]; }
		}
		for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) covar[(long)j * n + i] += P[(long)j * n + i] + P[(long)i * n + j];
		ap_valid_matrix(covar, &POP(s, sett, 0, 0, 0), n);
		update_pop(s, sett);
	}
	if (out_pop) memcpy(out_pop, s->pop, (size_t)(T + 1) * n * sizeof(double));
	free(covar); free(old); free(dtm); free(R); free(P);
	approx_free(x);
	return L;
}
