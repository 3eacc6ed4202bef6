/* beep_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the BEEP hot path.
 *
 * This file is the parity ORACLE for the CUDA engine: a plain, sequential C99
 * restatement of the reference algorithm (ScottishCovidResponse/BEEP @ d70908f)
 * over the flat BeepModelDesc arrays.  Every function cites the reference
 * file:line it follows.  It is pinned against the real reference (built from
 * /root/reference into oracle/_ref by oracle/ref/build_ref.sh) through the
 * golden vectors in tests/golden/ -- see tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  The product (beep_b200/) never does.
 *
 * Random numbers: the reference draws from global std:: distributions
 * (utils.cc:148-153,174-180).  Here and in the harness they are replaced by
 * the counter-based samplers of include/beep_rng.h, in one of two stream modes:
 *   SEQ    one sequential stream consumed in the reference's loop order
 *          (state.cc:614-631, mbp.cc:276-313) -- matches the reference harness;
 *   KEYED  one stream per (sett, area, transition, dp) cell
 *          (beep_stream_init_cell) -- matches the CUDA engine.
 * and the infectivity map is maintained either incrementally from transnum
 * (the reference, state.cc:164-219) or recomputed from the population each step
 * (the engine); both agree to rounding (state_check.cc:60-75 allows 1e-5).
 *
 * Compile with -ffp-contract=off (see oracle/Makefile).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

#include "beep_b200.h"
#include "beep_rng.h"

#define LARGE 1000000000.0  /* consts.hh:154 */
#define VTINY 0.000000000000001
#define TINY 0.00000001

typedef struct OracleState {
	const BeepModelDesc *m;
	int stochastic, keyed, imap_from_pop;
	int nR, nefoi, ninf, nmatmod, ncovar, nlevel, nsection;
	int *area_rgrp;             /* [A] index into Rspline_info                */
	int *section_ti, *section_tf, *section_off, *section_obs;
	/* State members, state.hh:16-63 */
	double *paramval, *paramv_dir;
	double *transrate, *disc_spline, *sus, *areafactor, *Ntime, *beta;
	double *pop, *transnum, *transmean, *Imap, *Idiag;
	double *eig_guess; int eig_n;
	double EF;
	long ndraw;
} OracleState;

static double *dalloc(long n) { double *p = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); return p; }

#define POP(s, sett, c, co, dp) (s)->pop[(((long)(sett) * m->A + (c)) * m->C + (co)) * m->D + (dp)]
#define TN(s, sett, c, tr, dp) (s)->transnum[(((long)(sett) * m->A + (c)) * m->J + (tr)) * m->D + (dp)]
#define TM(s, sett, c, tr, dp) (s)->transmean[(((long)(sett) * m->A + (c)) * m->J + (tr)) * m->D + (dp)]
#define IMAP(s, sett, st, v) (s)->Imap[((long)(sett) * m->S + (st)) * m->A * m->G + (v)]
#define IDIAG(s, sett, st, v) (s)->Idiag[((long)(sett) * m->S + (st)) * m->A * m->G + (v)]

OracleState *oracle_state_create(const BeepModelDesc *m)
{
	OracleState *s = (OracleState *)calloc(1, sizeof(OracleState));
	s->m = m;
	s->stochastic = m->stochastic; s->keyed = 0; s->imap_from_pop = 0;
	s->nR = (int)m->n_R_spline_ref; s->nefoi = (int)m->n_efoi_spline_ref; s->ninf = (int)m->n_inf_state;
	s->nmatmod = (int)m->n_matmod_spline; s->ncovar = (int)m->n_covar_param; s->nlevel = (int)m->n_level_param;
	s->area_rgrp = (int *)calloc(m->A, sizeof(int));
	for (int i = 0; i < s->nR; i++)
		for (int k = m->R_area_off[i]; k < m->R_area_off[i + 1]; k++) s->area_rgrp[m->R_area[k]] = i;
	long TAJD = (long)m->T * m->A * m->J * m->D;
	s->paramval = dalloc(m->nparam); s->paramv_dir = dalloc(m->nparam);
	s->transrate = dalloc((long)m->J * m->D); s->disc_spline = dalloc((long)m->nspline * m->T);
	s->sus = dalloc(m->D); s->areafactor = dalloc((long)m->period * m->A);
	s->Ntime = dalloc((long)m->T * m->G * m->G); s->beta = dalloc((long)m->S * m->A * m->T);
	s->pop = dalloc((long)(m->T + 1) * m->A * m->C * m->D);   /* pop[T] is used by the MVN-approximate likelihood only */
	s->transnum = dalloc(TAJD); s->transmean = dalloc(TAJD);
	s->Imap = dalloc((long)m->T * m->S * m->A * m->G); s->Idiag = dalloc((long)m->T * m->S * m->A * m->G);
	/* sections, obsmodel.cc:326-348 */
	int per = m->pmcmc_obs_period > 0 ? m->pmcmc_obs_period : m->T;
	s->nsection = (m->T + per - 1) / per;
	s->section_ti = (int *)calloc(s->nsection, sizeof(int)); s->section_tf = (int *)calloc(s->nsection, sizeof(int));
	s->section_off = (int *)calloc(s->nsection + 1, sizeof(int)); s->section_obs = (int *)calloc(m->nobs > 0 ? m->nobs : 1, sizeof(int));
	int cnt = 0;
	for (int sec = 0; sec < s->nsection; sec++) {
		int ti = sec * per, tf = ti + per; if (tf > m->T) tf = m->T;
		s->section_ti[sec] = ti; s->section_tf[sec] = tf; s->section_off[sec] = cnt;
		for (int i = 0; i < m->nobs; i++)
			if (m->obs_sett_i[i] < tf && m->obs_sett_f[i] >= ti && m->obs_sett_i[i] >= ti && m->obs_sett_f[i] <= tf) s->section_obs[cnt++] = i;
	}
	s->section_off[s->nsection] = cnt;
	return s;
}

void oracle_state_destroy(OracleState *s)
{
	if (!s) return;
	free(s->area_rgrp); free(s->section_ti); free(s->section_tf); free(s->section_off); free(s->section_obs);
	free(s->paramval); free(s->paramv_dir); free(s->transrate); free(s->disc_spline); free(s->sus); free(s->areafactor);
	free(s->Ntime); free(s->beta); free(s->pop); free(s->transnum); free(s->transmean); free(s->Imap); free(s->Idiag);
	free(s->eig_guess); free(s);
}

void oracle_set_mode(OracleState *s, int stochastic, int keyed, int imap_from_pop)
{
	s->stochastic = stochastic; s->keyed = keyed; s->imap_from_pop = imap_from_pop;
}

double *oracle_ptr(OracleState *s, int which)
{
	switch (which) {
	case 0: return s->transrate; case 1: return s->disc_spline; case 2: return s->sus; case 3: return s->areafactor;
	case 4: return s->beta; case 5: return s->Ntime; case 6: return s->pop; case 7: return s->transnum;
	case 8: return s->transmean; case 9: return s->Imap; case 10: return s->Idiag; case 11: return s->paramval;
	}
	return NULL;
}
int oracle_nsection(const OracleState *s) { return s->nsection; }
int oracle_section_range(const OracleState *s, int sec, int *ti, int *tf) { *ti = s->section_ti[sec]; *tf = s->section_tf[sec]; return 0; }
long oracle_ndraw(const OracleState *s) { return s->ndraw; }

/* ---------------------------------------------------------------------- */
/* K0  State::set_param, state.cc:66-91                                    */
/* ---------------------------------------------------------------------- */

/* matrix.cc:703-765 invert_matrix: Gauss-Jordan without pivoting */
static void invert_matrix(double *M, double *inv, int n)
{
	for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
	for (int ii = 0; ii < n; ii++) {
		double r = M[ii * n + ii];
		for (int i = ii; i < n; i++) M[ii * n + i] /= r;
		for (int i = 0; i <= ii; i++) inv[ii * n + i] /= r;
		for (int jj = ii + 1; jj < n; jj++) {
			double rr = M[jj * n + ii];
			if (rr != 0) {
				for (int i = ii; i < n; i++) M[jj * n + i] -= rr * M[ii * n + i];
				for (int i = 0; i <= ii; i++) inv[jj * n + i] -= rr * inv[ii * n + i];
			}
		}
	}
	for (int ii = n - 1; ii > 0; ii--)
		for (int jj = ii - 1; jj >= 0; jj--) {
			double rr = M[jj * n + ii];
			if (rr != 0) for (int i = 0; i < n; i++) inv[jj * n + i] -= rr * inv[ii * n + i];
		}
}

/* matrix.cc:1139-1220 largest_eigenvalue: power iteration, L1-normalised, warm
 * started along the time axis.  The reference seeds its first guess with ran()
 * and then carries it GLOBALLY across calls (matrix.cc:1136,1171-1173); here every
 * set_param starts from the flat vector (the harness resets the reference's global
 * to the same), so the result does not depend on the call history. */
static double largest_eigenvalue(OracleState *s, const double *M, int n)
{
	if (s->eig_n != n) {
		free(s->eig_guess); s->eig_guess = dalloc(n); s->eig_n = n;
		for (int i = 0; i < n; i++) s->eig_guess[i] = 1.0 / n;
	}
	double *vec = dalloc(n), *vec2 = dalloc(n);
	memcpy(vec, s->eig_guess, n * sizeof(double));
	double ev = 0;
	unsigned loop = 0;
	for (;;) {
		double sum = 0;
		for (int i = 0; i < n; i++) { double a = 0; for (int k = 0; k < n; k++) a += M[i * n + k] * vec[k]; vec2[i] = a; }
		for (int i = 0; i < n; i++) sum += vec2[i];
		double ev_new = sum;
		for (int i = 0; i < n; i++) vec2[i] /= sum;
		if (loop % 10 == 0) {
			double lim = 100 * VTINY;
			if (loop > 100) lim *= 10;
			if (loop > 1000) lim *= 10;
			if (ev_new - ev > -lim && ev_new - ev < lim) {
				int k;
				for (k = 0; k < n; k++) { double dif = vec2[k] - vec[k]; if (dif > lim || dif < -lim) break; }
				if (k == n) break;
			}
		}
		memcpy(vec, vec2, n * sizeof(double));
		ev = ev_new;
		loop++;
		if (loop > 10000) break;
	}
	memcpy(s->eig_guess, vec, n * sizeof(double));
	free(vec); free(vec2);
	return ev;
}

int oracle_set_param(OracleState *s, const double *theta)
{
	const BeepModelDesc *m = s->m;
	int T = m->T, A = m->A, J = m->J, D = m->D, G = m->G, S = m->S, Dq = m->Dq;
	memcpy(s->paramval, theta, m->nparam * sizeof(double));
	s->eig_n = 0;                                   /* flat start of the power iteration, see largest_eigenvalue */
	double *pd = s->paramv_dir;
	memcpy(pd, theta, m->nparam * sizeof(double));

	/* model.cc:2685-2695 dirichlet_correct */
	int ndir = (int)m->n_dir_off - 1;
	for (int g = 0; g < ndir; g++) {
		double sum = 0;
		for (int k = m->dir_off[g]; k < m->dir_off[g + 1]; k++) sum += theta[m->dir_th[k]];
		for (int k = m->dir_off[g]; k < m->dir_off[g + 1]; k++) pd[m->dir_th[k]] = theta[m->dir_th[k]] / (m->dir_frac[k] * sum);
	}

	/* model.cc:1116-1140 create_transrate */
	for (int tr = 0; tr < J; tr++)
		for (int dp = 0; dp < D; dp++) {
			double rate = 0;
			if (!m->trans_inf[tr]) {
				double mean = pd[m->trans_mean_param[tr * D + dp]] / m->trans_shape[tr];
				rate = 1.0 / mean;
				if (m->trans_prob_param[tr * D + dp] >= 0) rate *= pd[m->trans_prob_param[tr * D + dp]];
			}
			s->transrate[tr * D + dp] = rate;
		}

	/* model.cc:1020-1043 create_disc_spline */
	for (int sp = 0; sp < m->nspline; sp++) {
		double factor = pd[m->spline_factor_param[sp]];
		int o = m->spline_off[sp], np = m->spline_off[sp + 1] - o;
		int p = 0;
		for (int st = 0; st < T; st++) {
			double t = (double)((st + 0.5) * m->period) / T;
			while (p + 1 < np && t > m->spline_t[o + p + 1]) p++;
			double v;
			if (p + 1 == np) v = (st > 0) ? s->disc_spline[sp * T + st - 1] : 0;
			else {
				double fac = (t - m->spline_t[o + p]) / (m->spline_t[o + p + 1] - m->spline_t[o + p]);
				v = factor * (pd[m->spline_param[o + p]] * m->spline_multfac[o + p] * (1 - fac) +
				              pd[m->spline_param[o + p + 1]] * m->spline_multfac[o + p + 1] * fac);
			}
			s->disc_spline[sp * T + st] = v;
		}
		if (m->n_mod_spline_mult > 0)                                  /* model.cc:1044-1050: modifications of mode=prediction */
			for (int st = 0; st < T; st++) {
				s->disc_spline[sp * T + st] *= m->mod_spline_mult[sp * T + st];
				double set = m->mod_spline_set[sp * T + st];
				if (set != 999999999.0) s->disc_spline[sp * T + st] = set;
			}
	}

	/* model.cc:1144-1164 create_susceptibility */
	for (int dp = 0; dp < D; dp++) {
		double val = 1.0;
		for (int k = m->sus_off[dp]; k < m->sus_off[dp + 1]; k++) val *= pd[m->sus_param[k]];
		s->sus[dp] = val;
	}

	/* model.cc:1182-1250 create_areafactor + calculate_area_av; exp() is the shared beep_exp
	 * (include/beep_rng.h, ~1 ulp from libm) so that host and device agree bit for bit */
	{
		int period = m->period;
		double *af = dalloc(A);
		for (int c = 0; c < A; c++) {
			af[c] = (m->n_region_effect_param > 0) ? beep_exp(pd[m->region_effect_param[c]]) : 1;
			if (m->n_area_effect_param > 0) af[c] *= pd[m->area_effect_param[c]];
		}
		for (int t = 0; t < period; t++) for (int c = 0; c < A; c++) s->areafactor[t * A + c] = af[c];
		for (int j = 0; j < s->ncovar; j++)
			for (int t = 0; t < period; t++)
				for (int c = 0; c < A; c++)
					s->areafactor[t * A + c] *= beep_exp(pd[m->covar_param[j]] * m->covar_value[((long)j * A + c) * period + t]);
		if (s->nlevel > 0)
			for (int t = 0; t < period; t++)
				for (int c = 0; c < A; c++) s->areafactor[t * A + c] *= pd[m->level_param[m->level_map[t * A + c]]];
		double area_av = 0;
		for (int c = 0; c < A; c++) {
			double av = 0;
			for (int t = 0; t < period; t++) av += s->areafactor[t * A + c];
			av /= period;
			area_av += av * m->area_total_pop[c];
		}
		area_av /= m->popsize;
		for (int t = 0; t < period; t++) for (int c = 0; c < A; c++) s->areafactor[t * A + c] /= area_av;
		free(af);
	}

	/* model.cc:1254-1301 create_Ntime */
	for (int st = 0; st < T; st++) {
		double *N = s->Ntime + (long)st * G * G;
		memcpy(N, m->N0, (size_t)G * G * sizeof(double));
		for (int mm = 0; mm < s->nmatmod; mm++) {
			double fac = s->disc_spline[m->matmod_spline[mm] * T + st];
			for (int k = m->matmod_age_off[mm]; k < m->matmod_age_off[mm + 1]; k++) {
				int a = m->matmod_age[k];
				for (int i = 0; i < G; i++) { N[a * G + i] *= fac; N[i * G + a] *= fac; }
			}
		}
	}

	/* model.cc:1318-1362 calculate_beta_from_R */
	{
		int ninf = s->ninf, N = Dq * ninf;
		double *V = dalloc((long)N * N), *Vinv = dalloc((long)N * N), *F = dalloc((long)Dq * N), *NGM = dalloc((long)Dq * Dq);
		int *inf_ref = (int *)malloc(m->C * sizeof(int));
		for (int co = 0; co < m->C; co++) inf_ref[co] = -1;
		for (int i = 0; i < ninf; i++) inf_ref[m->inf_state[i]] = i;
		for (int st = 0; st < S; st++) {
			/* model.cc:2485-2522 calculate_Vinv */
			memset(V, 0, (size_t)N * N * sizeof(double));
			for (int tr = 0; tr < J; tr++) {
				int from = inf_ref[m->trans_from[tr]], to = inf_ref[m->trans_to[tr]];
				if (from >= 0)
					for (int dp = 0; dp < Dq; dp++) {
						double rate = s->transrate[tr * D + st * Dq + dp];
						int j = from * Dq + dp;
						V[j * N + j] += rate;
						if (to >= 0) { int k = to * Dq + dp; V[k * N + j] -= rate; }
					}
			}
			invert_matrix(V, Vinv, N);
			for (int ir = 0; ir < s->nR; ir++) {
				int spl = m->R_spline_ref[ir];
				int c0 = m->R_area[m->R_area_off[ir]];
				const double *dist = m->R_dist + (long)ir * Dq;
				double ratio = 0;
				for (int sett = 0; sett < T; sett++) {
					const double *Nt = s->Ntime + (long)sett * G * G;
					int recompute = (sett == 0);
					if (!recompute) { /* model.cc:1305-1314 equal() */
						const double *Np = Nt - (long)G * G;
						for (int i = 0; i < G * G; i++) if (Np[i] < Nt[i] - TINY || Np[i] > Nt[i] + TINY) { recompute = 1; break; }
					}
					if (recompute) {
						/* model.cc:2526-2556 calculate_F */
						memset(F, 0, (size_t)Dq * N * sizeof(double));
						for (int is = 0; is < ninf; is++) {
							double inf = pd[m->comp_inf_param[m->inf_state[is]]];
							if (inf != 0)
								for (int dp = 0; dp < Dq; dp++) {
									int i = is * Dq + dp, a = m->dp_age[dp];
									for (int q = 0; q < Dq; q++) {
										int aa = m->dp_age[q];
										F[q * N + i] += dist[q] * s->sus[Dq * st + q] * inf * Nt[aa * G + a];
									}
								}
						}
						/* model.cc:2560-2576 calculate_NGM */
						for (int q = 0; q < Dq; q++)
							for (int i = 0; i < Dq; i++) {
								double sum = 0;
								for (int j = 0; j < N; j++) sum += F[q * N + j] * Vinv[j * N + i];
								NGM[q * Dq + i] = sum;
							}
						ratio = largest_eigenvalue(s, NGM, Dq);
					}
					s->beta[((long)st * A + c0) * T + sett] = pd[m->strain_Rfactor_param[st]] * s->disc_spline[spl * T + sett] / ratio;
				}
				for (int k = m->R_area_off[ir] + 1; k < m->R_area_off[ir + 1]; k++)
					memcpy(s->beta + ((long)st * A + m->R_area[k]) * T, s->beta + ((long)st * A + c0) * T, T * sizeof(double));
			}
		}
		free(V); free(Vinv); free(F); free(NGM); free(inf_ref);
	}
	if (m->n_mod_beta > 0)                                             /* model.cc:1349-1357 */
		for (int st = 0; st < S; st++)
			for (int c = 0; c < A; c++)
				for (int sett = 0; sett < T; sett++) s->beta[((long)st * A + c) * T + sett] *= m->mod_beta[((long)sett * A + c) * S + st];
	return 0;
}

/* ---------------------------------------------------------------------- */
/* K1  infectivity map                                                      */
/* ---------------------------------------------------------------------- */

/* state.cc:266-278 / 204-216: spread dinf of area c through row c of M */
static void spread_dinf(const BeepModelDesc *m, double *Ima, double *Idia, int c, const double *dinf)
{
	int G = m->G;
	double diag = m->M_diag[c];
	int v = c * G;
	for (int a = 0; a < G; a++) {
		double di = dinf[a];
		if (di != 0) {
			Idia[v + a] += di * diag;
			for (int j = m->M_off[c]; j < m->M_off[c + 1]; j++) Ima[m->M_to[j] * G + a] += di * m->M_val[j];
		}
	}
}

/* state.cc:223-290 set_I_from_pop */
static void set_I_from_pop(OracleState *s, int sett)
{
	const BeepModelDesc *m = s->m;
	int A = m->A, G = m->G, Dq = m->Dq;
	double *dinf = dalloc(G);
	for (int st = 0; st < m->S; st++) {
		double *Ima = &IMAP(s, sett, st, 0), *Idia = &IDIAG(s, sett, st, 0);
		for (int v = 0; v < A * G; v++) { Ima[v] = 0; Idia[v] = 0; }
		for (int c = 0; c < A; c++) {
			for (int a = 0; a < G; a++) dinf[a] = 0;
			for (int co = 0; co < m->C; co++) {
				double inf = s->paramval[m->comp_inf_param[co]];
				if (inf != 0)
					for (int dp = 0; dp < Dq; dp++) {
						double p = POP(s, sett, c, co, st * Dq + dp);
						if (p != 0) dinf[m->dp_age[dp]] += inf * p;
					}
			}
			spread_dinf(m, Ima, Idia, c, dinf);
		}
	}
	free(dinf);
}

/* state.cc:178-219 update_I_from_transnum; dtn is [A][J][D] */
static void update_I_from_transnum(OracleState *s, double *Ima_all, double *Idia_all, const double *dtn)
{
	const BeepModelDesc *m = s->m;
	int A = m->A, G = m->G, Dq = m->Dq, J = m->J, D = m->D;
	double *dinf = dalloc(G);
	for (int st = 0; st < m->S; st++) {
		double *Ima = Ima_all + (long)st * A * G, *Idia = Idia_all + (long)st * A * G;
		for (int c = 0; c < A; c++) {
			for (int a = 0; a < G; a++) dinf[a] = 0;
			for (int tr = 0; tr < J; tr++) {
				/* model.cc:820-823 get_infectivity_dif */
				double di = s->paramval[m->comp_inf_param[m->trans_to[tr]]] - s->paramval[m->comp_inf_param[m->trans_from[tr]]];
				if (di != 0)
					for (int dp = 0; dp < Dq; dp++) {
						double num = dtn[((long)c * J + tr) * D + st * Dq + dp];
						if (num != 0) dinf[m->dp_age[dp]] += di * num;
					}
			}
			spread_dinf(m, Ima, Idia, c, dinf);
		}
	}
	free(dinf);
}

/* state.cc:164-174 set_Imap_sett */
static void set_Imap_sett(OracleState *s, int sett)
{
	const BeepModelDesc *m = s->m;
	if (sett == 0) { set_I_from_pop(s, 0); return; }
	long n = (long)m->S * m->A * m->G;
	memcpy(&IMAP(s, sett, 0, 0), &IMAP(s, sett - 1, 0, 0), n * sizeof(double));
	memcpy(&IDIAG(s, sett, 0, 0), &IDIAG(s, sett - 1, 0, 0), n * sizeof(double));
	update_I_from_transnum(s, &IMAP(s, sett, 0, 0), &IDIAG(s, sett, 0, 0), &TN(s, sett - 1, 0, 0, 0));
}

/* ---------------------------------------------------------------------- */
/* K2  transition means                                                     */
/* ---------------------------------------------------------------------- */

/* state.cc:547-578 get_NMI */
static void get_NMI(OracleState *s, int sett, int st, int c, double *NMI)
{
	const BeepModelDesc *m = s->m;
	int G = m->G;
	double d = s->disc_spline[m->geo_spline_ref * m->T + sett];
	double omd = 1 - d;
	double fac = m->M_factor[c];
	if (G == 1) {
		double val = IDIAG(s, sett, st, c) * (d + fac * omd) + IMAP(s, sett, st, c) * d;
		if (val < 0) val = 0;
		NMI[0] = val;
		return;
	}
	double *I = dalloc(G);
	int v = c * G;
	for (int a = 0; a < G; a++) I[a] = (IDIAG(s, sett, st, v + a) + IMAP(s, sett, st, v + a)) * d + fac * IDIAG(s, sett, st, v + a) * omd;
	const double *Nt = s->Ntime + (long)sett * G * G;
	for (int a = 0; a < G; a++) {
		double sum = 0;
		for (int aa = 0; aa < G; aa++) sum += Nt[a * G + aa] * I[aa];
		if (sum < 0) sum = 0;
		NMI[a] = sum;
	}
	free(I);
}

/* state.cc:380-456 set_transmean */
static void set_transmean(OracleState *s, int sett, int c)
{
	const BeepModelDesc *m = s->m;
	int G = m->G, Dq = m->Dq, D = m->D, S = m->S, J = m->J, T = m->T, A = m->A;
	double dt = (double)m->period / T;
	int itr = m->infection_trans;
	int sfrom = m->trans_from[itr];
	if (S > 1) { /* state.cc:395-404: susceptible pool shift, mutates pop */
		for (int dp = 0; dp < Dq; dp++) {
			double sum = 0;
			for (int st = 1; st < S; st++) { sum += POP(s, sett, c, sfrom, Dq * st + dp); POP(s, sett, c, sfrom, Dq * st + dp) = 0; }
			POP(s, sett, c, sfrom, dp) += sum;
		}
	}
	double *NMI = dalloc(G), *eta_age = dalloc(G);
	for (int st = 0; st < S; st++) {
		get_NMI(s, sett, st, c, NMI);
		double be = s->beta[((long)st * A + c) * T + sett];
		be *= s->areafactor[(sett / m->spu) * A + c];
		int ie = m->efoi_ref[st * A + c];
		double et = s->disc_spline[m->efoi_spline_ref[ie] * T + sett] / m->popsize;
		if (m->n_mod_efoi > 0) et *= m->mod_efoi[((long)sett * A + c) * S + st];            /* state.cc:415 */
		for (int a = 0; a < G; a++) eta_age[a] = et * m->efoi_agedist[ie * G + a];
		for (int dp = 0; dp < Dq; dp++) {
			int dpp = st * Dq + dp;
			double popu = POP(s, sett, c, sfrom, dp);
			if (popu <= 0) TM(s, sett, c, itr, dpp) = 0;
			else {
				int a = m->dp_age[dp];
				TM(s, sett, c, itr, dpp) = dt * popu * s->sus[dpp] * (be * NMI[a] + eta_age[a]);
			}
		}
	}
	for (int tr = 0; tr < J; tr++)
		if (!m->trans_inf[tr]) {
			int from = m->trans_from[tr];
			for (int dp = 0; dp < D; dp++) {
				double popu = POP(s, sett, c, from, dp);
				if (popu <= 0) TM(s, sett, c, tr, dp) = 0;
				else TM(s, sett, c, tr, dp) = dt * popu * s->transrate[tr * D + dp];
			}
		}
	if (m->n_mod_transmean > 0)                                        /* state.cc:446-453 */
		for (int tr = 0; tr < J; tr++)
			for (int dp = 0; dp < D; dp++) TM(s, sett, c, tr, dp) *= m->mod_transmean[((((long)sett * A + c) * J) + tr) * D + dp];
	free(NMI); free(eta_age);
}

/* state.cc:460-485 update_pop */
static void update_pop(OracleState *s, int sett)
{
	const BeepModelDesc *m = s->m;
	long n = (long)m->A * m->C * m->D;
	memcpy(&POP(s, sett + 1, 0, 0, 0), &POP(s, sett, 0, 0, 0), n * sizeof(double));
	for (int c = 0; c < m->A; c++)
		for (int tr = 0; tr < m->J; tr++) {
			int from = m->trans_from[tr], to = m->trans_to[tr];
			for (int dp = 0; dp < m->D; dp++) {
				double num = TN(s, sett, c, tr, dp);
				if (num != 0) { POP(s, sett + 1, c, from, dp) -= num; POP(s, sett + 1, c, to, dp) += num; }
			}
		}
}

/* state.cc:339-376 democat_change_pop_adjust: the population of the areas of a demographic change is shifted between
 * the categories of its classification so that their fractions follow the data (same loops, same order of sums) */
static void democat_change_pop_adjust(OracleState *s, int sett)
{
	const BeepModelDesc *m = s->m;
	const int ndcc = (int)(m->n_dcc_info / 6);
	for (int k = 0; k < ndcc; k++) {
		const int32_t *inf = m->dcc_info + 6 * k;
		const int ncat = inf[0], na = inf[2], ng = inf[3];
		const int32_t *area = m->dcc_area + inf[1], *grp = m->dcc_group + inf[4];
		const double *frac = m->dcc_frac + inf[5] + (long)sett * ncat;
		double pop_cat[64], frac_dif[64];
		for (int f = 0; f < ncat; f++) pop_cat[f] = 0;
		for (int ia = 0; ia < na; ia++)
			for (int g = 0; g < ng; g++)
				for (int co = 0; co < m->C; co++)
					for (int f = 0; f < ncat; f++) pop_cat[f] += POP(s, sett, area[ia], co, grp[g * ncat + f]);
		double pop_tot = 0.0;
		for (int f = 0; f < ncat; f++) pop_tot += pop_cat[f];
		for (int f = 0; f < ncat; f++) frac_dif[f] = frac[f] - pop_cat[f] / pop_tot;
		for (int ia = 0; ia < na; ia++)
			for (int g = 0; g < ng; g++)
				for (int co = 0; co < m->C; co++) {
					double sum = 0.0;
					for (int f = 0; f < ncat; f++) sum += POP(s, sett, area[ia], co, grp[g * ncat + f]);
					for (int f = 0; f < ncat; f++) POP(s, sett, area[ia], co, grp[g * ncat + f]) += frac_dif[f] * sum;
				}
	}
}

static void pop_init(OracleState *s)
{
	const BeepModelDesc *m = s->m;
	memcpy(s->pop, m->pop_init, (size_t)m->A * m->C * m->D * sizeof(double));
}

static uint64_t cell_id(const BeepModelDesc *m, int sett, int c, int tr, int dp) { return (uint64_t)((((long)sett * m->A + c) * m->J + tr) * m->D + dp); }

/* ---------------------------------------------------------------------- */
/* K1+K2  State::simulate(ti,tf), state.cc:596-638                          */
/* ---------------------------------------------------------------------- */
int oracle_simulate(OracleState *s, int ti, int tf, uint64_t seed, uint64_t stream_id)
{
	const BeepModelDesc *m = s->m;
	int J = m->J, D = m->D;
	BeepStream seq, cell;
	/* the sequential stream restarts at every call, as the harness re-keys it */
	beep_stream_init(&seq, seed, (uint32_t)stream_id, (uint32_t)(stream_id >> 32), 0xFFFFFFFFu);
	if (ti == 0) pop_init(s);
	for (int sett = ti; sett < tf; sett++) {
		democat_change_pop_adjust(s, sett);                 /* state.cc:607 */
		if (sett == ti || s->imap_from_pop) set_I_from_pop(s, sett);
		else set_Imap_sett(s, sett);
		for (int c = 0; c < m->A; c++) {
			set_transmean(s, sett, c);
			for (int tr = 0; tr < J; tr++)
				for (int dp = 0; dp < D; dp++) {
					double mean = TM(s, sett, c, tr, dp);
					if (mean == 0) TN(s, sett, c, tr, dp) = 0;
					else if (s->stochastic) {
						if (s->keyed) beep_stream_init_cell(&cell, seed, stream_id, cell_id(m, sett, c, tr, dp));
						TN(s, sett, c, tr, dp) = beep_poisson(s->keyed ? &cell : &seq, mean); s->ndraw++;
					}
					else TN(s, sett, c, tr, dp) = mean;
				}
		}
		if (sett < m->T - 1) update_pop(s, sett);
	}
	return 0;
}

/* state.cc:294-322 initialise_from_particle (set_param must have been called) */
int oracle_init_from_particle(OracleState *s, const double *transnum)
{
	const BeepModelDesc *m = s->m;
	int T = m->T;
	if (transnum != s->transnum) memcpy(s->transnum, transnum, (size_t)T * m->A * m->J * m->D * sizeof(double));
	pop_init(s);
	for (int sett = 0; sett < T; sett++) { democat_change_pop_adjust(s, sett); if (sett < T - 1) update_pop(s, sett); }   /* state.cc:307-310 */
	if (s->imap_from_pop) for (int sett = 0; sett < T; sett++) set_I_from_pop(s, sett);
	else { /* state.cc:127-160 set_Imap */
		set_I_from_pop(s, 0);
		long n = (long)m->S * m->A * m->G;
		double *Ima = dalloc(n), *Idia = dalloc(n);
		memcpy(Ima, &IMAP(s, 0, 0, 0), n * sizeof(double)); memcpy(Idia, &IDIAG(s, 0, 0, 0), n * sizeof(double));
		for (int sett = 0; sett < T; sett++) {
			memcpy(&IMAP(s, sett, 0, 0), Ima, n * sizeof(double)); memcpy(&IDIAG(s, sett, 0, 0), Idia, n * sizeof(double));
			update_I_from_transnum(s, Ima, Idia, &TN(s, sett, 0, 0, 0));
		}
		free(Ima); free(Idia);
	}
	for (int sett = 0; sett < T; sett++) for (int c = 0; c < m->A; c++) set_transmean(s, sett, c);
	return 0;
}

/* mbp.cc:470-476: set_transmean for every (sett, c) with the current parameters,
 * populations and infectivity maps left as they are */
int oracle_recompute_transmean(OracleState *s)
{
	const BeepModelDesc *m = s->m;
	for (int sett = 0; sett < m->T; sett++) for (int c = 0; c < m->A; c++) set_transmean(s, sett, c);
	return 0;
}

/* copy of another state's latent arrays (Mbp::copy_initial_to_propse) */
int oracle_copy_state(OracleState *dst, const OracleState *src)
{
	const BeepModelDesc *m = src->m;
	long TAJD = (long)m->T * m->A * m->J * m->D;
	memcpy(dst->paramval, src->paramval, m->nparam * sizeof(double));
	memcpy(dst->pop, src->pop, (size_t)m->T * m->A * m->C * m->D * sizeof(double));
	memcpy(dst->transnum, src->transnum, TAJD * sizeof(double));
	memcpy(dst->transmean, src->transmean, TAJD * sizeof(double));
	memcpy(dst->Imap, src->Imap, (size_t)m->T * m->S * m->A * m->G * sizeof(double));
	memcpy(dst->Idiag, src->Idiag, (size_t)m->T * m->S * m->A * m->G * sizeof(double));
	return 0;
}

/* ---------------------------------------------------------------------- */
/* K3  Mbp::mbp, mbp.cc:252-330 (prop must have had oracle_set_param)       */
/* simu_mask: [T][A] bytes, 1 = SIMU, or NULL                               */
/* ---------------------------------------------------------------------- */
int oracle_mbp(const OracleState *init, OracleState *prop, const unsigned char *simu_mask, uint64_t seed, uint64_t stream_id)
{
	OracleState *s = prop;
	const BeepModelDesc *m = s->m;
	int A = m->A, J = m->J, D = m->D, T = m->T;
	long nI = (long)m->S * A * m->G;
	double *dImap = dalloc(nI), *dIdiag = dalloc(nI), *dtn = dalloc((long)A * J * D);
	BeepStream seq, cell;
	beep_stream_init(&seq, seed, (uint32_t)stream_id, (uint32_t)(stream_id >> 32), 0xFFFFFFFFu);
	pop_init(s); /* mbp.cc:414-424 mbp_initialise */
	for (int sett = 0; sett < T; sett++) {
		democat_change_pop_adjust(s, sett);                 /* mbp.cc:265 */
		if (s->imap_from_pop) set_I_from_pop(s, sett);
		else { /* INF_DIF_UPDATE, state.cc:109-123 set_Imap_using_dI */
			for (long v = 0; v < nI; v++) {
				(&IMAP(s, sett, 0, 0))[v] = (&IMAP(init, sett, 0, 0))[v] + dImap[v];
				(&IDIAG(s, sett, 0, 0))[v] = (&IDIAG(init, sett, 0, 0))[v] + dIdiag[v];
			}
		}
		for (int c = 0; c < A; c++) {
			set_transmean(s, sett, c);
			int simu = simu_mask ? simu_mask[(long)sett * A + c] : 0;
			for (int tr = 0; tr < J; tr++)
				for (int dp = 0; dp < D; dp++) {
					BeepStream *rs = &seq;
					if (s->keyed) { rs = &cell; beep_stream_init_cell(rs, seed, stream_id, cell_id(m, sett, c, tr, dp)); }
					int num_i = (int)TN(init, sett, c, tr, dp);
					double rate_p = TM(s, sett, c, tr, dp);
					int num_p;
					if (rate_p == 0) num_p = 0;
					else if (simu) { num_p = beep_poisson(rs, rate_p); s->ndraw++; }
					else {
						double rate_i = TM(init, sett, c, tr, dp);
						if (rate_p > rate_i) { num_p = num_i + beep_poisson(rs, rate_p - rate_i); s->ndraw++; }
						else if (rate_p == rate_i) num_p = num_i;
						else if (num_i == 0) num_p = 0; /* utils.cc:176 */
						else { num_p = beep_binomial(rs, rate_p / rate_i, num_i); s->ndraw++; }
					}
					dtn[((long)c * J + tr) * D + dp] = num_p - num_i;
					TN(s, sett, c, tr, dp) = num_p;
				}
		}
		if (sett < T - 1) {
			update_pop(s, sett);
			if (!s->imap_from_pop) update_I_from_transnum(s, dImap, dIdiag, dtn);
		}
	}
	free(dImap); free(dIdiag); free(dtn);
	return 0;
}

/* ---------------------------------------------------------------------- */
/* K4  observation model, obsmodel.cc                                       */
/* ---------------------------------------------------------------------- */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* obsmodel.cc:74-109 get_obs_value_section restated obs-major: the reference
 * visits cells in (sett, c, tr, dp) then (co, dp) order and adds to every
 * observation listed under the cell, so each observation receives its terms in
 * exactly the nesting below. */
static double obs_value_one(const OracleState *s, int ob)
{
	const BeepModelDesc *m = s->m;
	int tab = m->obs_table[ob];
	int na = m->obs_area_off[ob + 1] - m->obs_area_off[ob], nd = m->obs_dp_off[ob + 1] - m->obs_dp_off[ob];
	int nt = m->table_trans_off[tab + 1] - m->table_trans_off[tab], nc = m->table_comp_off[tab + 1] - m->table_comp_off[tab];
	int *ar = (int *)malloc((na + nd + nt + nc + 4) * sizeof(int)), *dp = ar + na, *tl = dp + nd, *cl = tl + nt;
	memcpy(ar, m->obs_area + m->obs_area_off[ob], na * sizeof(int)); qsort(ar, na, sizeof(int), cmp_int);
	memcpy(dp, m->obs_dp + m->obs_dp_off[ob], nd * sizeof(int)); qsort(dp, nd, sizeof(int), cmp_int);
	memcpy(tl, m->table_trans + m->table_trans_off[tab], nt * sizeof(int)); qsort(tl, nt, sizeof(int), cmp_int);
	memcpy(cl, m->table_comp + m->table_comp_off[tab], nc * sizeof(int)); qsort(cl, nc, sizeof(int), cmp_int);
	double val = 0, factor = m->obs_factor[ob];
	int spl = m->obs_factor_spline[ob];
	for (int sett = m->obs_sett_i[ob]; sett < m->obs_sett_f[ob]; sett++)
		for (int ia = 0; ia < na; ia++) {
			int c = ar[ia];
			for (int it = 0; it < nt; it++)
				for (int id = 0; id < nd; id++) {
					double num = TN(s, sett, c, tl[it], dp[id]);
					if (num != 0) {
						if (spl < 0) val += factor * num;
						else val += factor * num * s->disc_spline[spl * m->T + sett];
					}
				}
			for (int ic = 0; ic < nc; ic++)
				for (int id = 0; id < nd; id++) val += factor * POP(s, sett, c, cl[ic], dp[id]);
		}
	free(ar);
	return val;
}

void oracle_obs_values(const OracleState *s, double *out)
{
	for (int ob = 0; ob < s->m->nobs; ob++) out[ob] = obs_value_one(s, ob);
}

static double normal_probability(double x, double mean, double var) { return -0.5 * log(2 * M_PI * var) - (x - mean) * (x - mean) / (2 * var); } /* utils.cc:107-111 */
static double poisson_probability(int i, double lam) /* utils.cc:157-170 */
{
	if (lam == 0) return (i == 0) ? 0 : -LARGE;
	if (i < 0) return -LARGE;
	return i * log(lam) - lam - lgamma(i + 1);
}
static double negative_binomial_probability(unsigned val, double mm, double shape) /* utils.cc:124-127 */
{
	return lgamma(shape + val) - lgamma(val + 1) - lgamma(shape) + shape * log(shape / (shape + mm)) + val * log(mm / (shape + mm));
}

/* obsmodel.cc:245-305 obs_prob */
double oracle_obs_prob(const OracleState *s, double value, int ob)
{
	const BeepModelDesc *m = s->m;
	if (value < 0) value = 0;
	double val = m->obs_value[ob], sd = m->obs_sd[ob];
	if (val == BEEP_UNKNOWN) return 0;
	int om = m->obs_model[ob];
	if (val == BEEP_THRESH) {
		double thresh = m->table_threshold[m->obs_table[ob]];
		if (om == BEEP_OBS_NORMAL || om == BEEP_OBS_NORMAL_PERCENT) {
			if (value < thresh) return normal_probability(thresh, thresh, sd * sd);
			return normal_probability(thresh, value, sd * sd);
		}
		if (om == BEEP_OBS_POISSON) {
			double lam = value; if (lam < 1) lam = 1;
			double sum = 0; for (double v = 0; v <= thresh; v++) sum += exp(poisson_probability((int)v, lam));
			return log(sum);
		}
		double mm = value; if (mm < 1) mm = 1;
		double sum = 0; for (double v = 0; v <= thresh; v++) sum += exp(negative_binomial_probability((unsigned)v, mm, m->obs_shape[ob]));
		return log(sum);
	}
	if (om == BEEP_OBS_NORMAL || om == BEEP_OBS_NORMAL_PERCENT) return normal_probability(val, value, sd * sd);
	if (om == BEEP_OBS_POISSON) { double lam = value; if (lam < 1) lam = 1; return poisson_probability((int)(val + 0.5), lam); }
	double mm = value; if (mm < 1) mm = 1;
	return negative_binomial_probability((unsigned)(int)(val + 0.5), mm, m->obs_shape[ob]);
}

/* obsmodel.cc:22-35 calculate */
double oracle_obs_loglik(const OracleState *s)
{
	double L = 0;
	for (int ob = 0; ob < s->m->nobs; ob++) L += oracle_obs_prob(s, obs_value_one(s, ob), ob);
	return L;
}

/* obsmodel.cc:39-58 calculate_section */
double oracle_obs_loglik_section(const OracleState *s, int sec)
{
	double L = 0;
	for (int k = s->section_off[sec]; k < s->section_off[sec + 1]; k++) { int ob = s->section_obs[k]; L += oracle_obs_prob(s, obs_value_one(s, ob), ob); }
	return L;
}

/* obsmodel.cc:309-322 get_EF_datatable */
void oracle_EF_datatable(const OracleState *s, double *out)
{
	for (int i = 0; i < s->m->ntable; i++) out[i] = 0;
	for (int ob = 0; ob < s->m->nobs; ob++) out[s->m->obs_table[ob]] += -2 * oracle_obs_prob(s, obs_value_one(s, ob), ob);
}

/* ---------------------------------------------------------------------- */
/* K5  State::likelihood, state.cc:812-826                                  */
/* ---------------------------------------------------------------------- */
double oracle_latent_loglik(const OracleState *s)
{
	const BeepModelDesc *m = s->m;
	long n = (long)m->T * m->A * m->J * m->D;
	double Li = 0;
	for (long i = 0; i < n; i++) Li += poisson_probability((int)s->transnum[i], s->transmean[i]);
	return Li;
}

/* ---------------------------------------------------------------------- */
/* PMCMC bootstrap, pmcmc.cc:211-268, with the engine's keyed uniform       */
/* stream (seed, section) instead of the global ran()                       */
/* ---------------------------------------------------------------------- */
int oracle_bootstrap(const double *L, int ntot, double invT, uint64_t seed, int section, double *loglik_inc, uint32_t *backpart)
{
	double Lmax = -1e300;
	for (int i = 0; i < ntot; i++) if (L[i] > Lmax) Lmax = L[i];
	double *w = dalloc(ntot), *wsum = dalloc(ntot);
	double sum = 0;
	for (int i = 0; i < ntot; i++) { w[i] = exp(invT * (L[i] - Lmax)); sum += w[i]; wsum[i] = sum; }
	*loglik_inc = invT * Lmax + log(sum / ntot);
	BeepStream rs; beep_stream_init(&rs, seed, (uint32_t)section, 0xB0075u, 0xFFFFFFFEu);
	unsigned char *flag = (unsigned char *)calloc(ntot, 1);
	uint32_t *extra = (uint32_t *)malloc(ntot * sizeof(uint32_t)); int nextra = 0;
	for (int i = 0; i < ntot; i++) backpart[i] = 0xFFFFFFFFu;
	for (int i = 0; i < ntot; i++) {
		double z = beep_next_uniform(&rs) * sum;
		int lo = 0, hi = ntot - 1;
		while (lo < hi) { int mid = (lo + hi) / 2; if (z <= wsum[mid]) hi = mid; else lo = mid + 1; }
		if (!flag[lo]) { flag[lo] = 1; backpart[lo] = (uint32_t)lo; } /* stay in place if chosen, pmcmc.cc:250-258 */
		else extra[nextra++] = (uint32_t)lo;
	}
	for (int i = 0; i < ntot; i++) if (backpart[i] == 0xFFFFFFFFu) backpart[i] = extra[--nextra];      /* pmcmc.cc:256-260: handed out from the back */
	free(w); free(wsum); free(flag); free(extra);
	return 0;
}

/* ---------------------------------------------------------------------- */
/* sampler parity helpers (host side of tests/test_rng_parity.py)           */
/* ---------------------------------------------------------------------- */
void oracle_poisson_many(uint64_t seed, uint64_t stream_id, const double *lam, int n, int32_t *out)
{
	for (int i = 0; i < n; i++) {
		BeepStream rs; beep_stream_init(&rs, seed, (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)i);
		out[i] = beep_poisson(&rs, lam[i]);
	}
}
void oracle_binomial_many(uint64_t seed, uint64_t stream_id, const double *p, const int32_t *nn, int n, int32_t *out)
{
	for (int i = 0; i < n; i++) {
		BeepStream rs; beep_stream_init(&rs, seed, (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)i);
		out[i] = beep_binomial(&rs, p[i], nn[i]);
	}
}
/* the exact inversion search from a given uniform (tests/test_rng_math_cpu.py: the device's single-precision pre-search) */
void oracle_poisson_inv_many(const double *u, const double *lam, int n, int32_t *out)
{
	for (int i = 0; i < n; i++) out[i] = beep_poisson_inv_u(u[i], lam[i]);
}
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { beep_philox4x32_10(ctr, key, out); }

/* the shared software exp / log (include/beep_rng.h), exposed so that tests can pin their accuracy against libm */
void oracle_exp_many(const double *x, int n, double *out) { for (int i = 0; i < n; i++) out[i] = beep_exp(x[i]); }
void oracle_log_many(const double *x, int n, double *out) { for (int i = 0; i < n; i++) out[i] = beep_log(x[i]); }

/* ---------------------------------------------------------------------- */
/* ABC-SMC importance-weight denominators            */
/* ---------------------------------------------------------------------- */
/* MVN::probability, mvn.cc:185-196: -1/2 SUM_v1 SUM_v2 (d_v1/size) inv_M[v1][v2] (d_v2/size), in the
 * reference's loop order */
static double oracle_mvn_probability(int nvar, const int32_t *var, const double *inv_M, double size, const double *pend, const double *pstart)
{
	double sum = 0;
	for (int v1 = 0; v1 < nvar; v1++)
		for (int v2 = 0; v2 < nvar; v2++) {
			double val1 = pend[var[v1]] - pstart[var[v1]];
			double val2 = pend[var[v2]] - pstart[var[v2]];
			sum += (val1 / size) * inv_M[v1 * nvar + v2] * (val2 / size);
		}
	return -0.5 * sum;
}
/* the sum of ABCSMC::calculate_particle_weight, abcsmc.cc:264-268, for each of N proposals:
 * out[i] = SUM_j w_last[j] exp(probability(prop_i, last_j)); the caller divides exp(prior) by it (:270) */
void oracle_mvn_kernel_sums(int nvar, int nparam, const int32_t *var, const double *inv_M, double size,
                            const double *param_prop, int N, const double *param_last, const double *w_last, int Nlast, double *out)
{
	for (int i = 0; i < N; i++) {
		double sum = 0;
		for (int j = 0; j < Nlast; j++)
			sum += w_last[j] * exp(oracle_mvn_probability(nvar, var, inv_M, size, param_prop + (size_t)i * nparam, param_last + (size_t)j * nparam));
		out[i] = sum;
	}
}

/* ---------------------------------------------------------------------- */
/* ObservationModel::get_graph_state, obsmodel.cc:113-170: the series behind */
/* the Posterior/state files, one block per graph                            */
/* ---------------------------------------------------------------------- */
int oracle_graph_npoints(const BeepModelDesc *m, int g)
{
	if (m->graph_type[g] == 1) return m->graph_point_off[g + 1] - m->graph_point_off[g];
	int gs = m->graph_step[0], n = 0;
	for (int sett = 0; sett < m->T - gs; sett += gs) n++;
	return n;
}

void oracle_graph_state(const OracleState *s, double *out)
{
	const BeepModelDesc *m = s->m;
	int ng = (int)m->n_graph_type, T = m->T, gs = m->graph_step[0];
	double *obs_value = dalloc(m->nobs);
	oracle_obs_values(s, obs_value);
	long k = 0;
	for (int g = 0; g < ng; g++) {
		int dt = m->graph_table[g], spl = m->graph_factor_spline[g];
		if (m->graph_type[g] == 0) {
			double fac_trans = m->graph_factor[g] * (double)m->spu / gs;
			double fac_pop = m->graph_factor[g] * 1.0 / gs;
			for (int sett = 0; sett < T - gs; sett += gs) {
				double sum = 0.0;
				for (int step = 0; step < gs; step++) {
					int sett2 = sett + step;
					for (int ia = m->graph_area_off[g]; ia < m->graph_area_off[g + 1]; ia++)
						for (int id = m->graph_dp_off[g]; id < m->graph_dp_off[g + 1]; id++) {
							int c = m->graph_area[ia], dp = m->graph_dp[id];
							for (int q = m->table_trans_off[dt]; q < m->table_trans_off[dt + 1]; q++) {
								double num = fac_trans * TN(s, sett2, c, m->table_trans[q], dp);
								if (spl < 0) sum += num; else sum += num * s->disc_spline[spl * T + sett2];
							}
							for (int q = m->table_comp_off[dt]; q < m->table_comp_off[dt + 1]; q++) {
								double num = fac_pop * POP(s, sett2, c, m->table_comp[q], dp);
								if (spl < 0) sum += num; else sum += num * s->disc_spline[spl * T + sett2];
							}
						}
				}
				out[k++] = sum;
			}
		} else {
			for (int q = m->graph_point_off[g]; q < m->graph_point_off[g + 1]; q++) out[k++] = obs_value[m->graph_point_obs[q]];
		}
	}
	free(obs_value);
}

/* ---------------------------------------------------------------------- */
/* MVN-approximate likelihood (mode=map), state_mvn_approx.cc               */
/* ---------------------------------------------------------------------- */
#include "beep_oracle_approx.c"
