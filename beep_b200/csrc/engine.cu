/* engine.cu -- C-ABI of the B200-native BEEP latent-state engine (include/beep_b200.h).
 *
 * Host logic only: validates and uploads the flat model description, builds the
 * derived read-only indices, owns batches of particles in HBM and launches the
 * kernels of kernels.cuh.  There is no CPU fallback: without a CUDA device
 * beep_engine_create fails with BEEP_ERR_NO_DEVICE.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "traj_chunk.cuh"
#include "imap_dense.cuh"
#include "approx.cuh"
#include "graphs.cuh"

/* ------------------------------------------------------------------------ */
/* errors                                                                    */
/* ------------------------------------------------------------------------ */
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	g_err = buf;
	return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(BEEP_ERR_CUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char *beep_last_error(void) { return g_err.c_str(); }
extern "C" int beep_version(void) { return 100; }
extern "C" int beep_desc_size(void) { return (int)sizeof(BeepModelDesc); }

/* ------------------------------------------------------------------------ */
/* engine                                                                    */
/* ------------------------------------------------------------------------ */
struct BeepEngine {
	int device = 0, nsm = 0, smem_optin = 0;
	DevModel dm;
	std::vector<void *> dev_allocs;
	std::vector<int> section_ti, section_tf, h_rgrp;
	unsigned char h_from[32] = {0}, h_to[32] = {0}, h_kind[32] = {0};
	long long launches = 0;
	int sync_mode = SYNC_NONE, nwarp = 4;
	size_t smem_traj = 0;
	int coop_grid = 0;
	int none_resident = 1 << 30;   /* CTAs of the tile-resident flavour that are resident at once */
	void *mvn_buf = nullptr; size_t mvn_cap = 0;   /* grow-only scratch of beep_mvn_kernel_sums */
	/* chunk kernel (traj_chunk.cuh): the model qualifies, tiles per chunk, launch shape */
	bool dense_m = false;              /* the mixing matrix is dense: Imap of a step by one DMMA GEMM for the whole batch (imap_dense.cuh) */
	bool chunk_model = false, chunk_ok = false;
	int chunk_nt = 0, chunk_grid = 0;
	size_t chunk_smem = 0;
	ChunkSm chunk_ls;
	/* MVN-approximate likelihood (approx.cuh): index lists built at creation when the model qualifies */
	bool approx_ok = false; std::string approx_why;
	DevApprox ax;
	/* state graphs (graphs.cuh) */
	bool graphs_ok = false, graphs_pop = false, graphs_obs = false;
	DevGraphs gx;
	std::vector<int> graph_point_off;
	std::vector<double> h_mod_beta;    /* host copy: beep_get_derived expands beta to the reference's [S][A][T] */
};

struct BeepBatch {
	BeepEngine *eng = nullptr;
	DevBatch db;
	std::vector<void *> dev_allocs;
	long long bytes = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	bool have_params = false, have_sim_time = false;
	std::vector<unsigned char> slot_theta;   /* slots whose parameters were given (set_params: all; import_particle: one) */
	bool theta_dirty = false;       /* a slot's parameters were replaced (beep_import_particle): the derived tables are re-derived before the next launch that reads them */
	int cur_t = 0;
	double *d_scratch = nullptr;    /* [P * max(k5_chunks, OBS_CHUNKS*(1+ntable))] */
	double *h_pinned = nullptr; size_t h_pinned_n = 0;
	unsigned *d_map = nullptr; unsigned char *d_flag = nullptr;
	unsigned char *d_mask = nullptr; size_t d_mask_n = 0;
	double *d_tmp = nullptr; size_t d_tmp_n = 0;
	double *d_approx = nullptr; size_t d_approx_n = 0;   /* scratch of beep_likelihood_approx */
};

template <typename Tp>
static int upload(BeepEngine *eng, const Tp *src, size_t n, const Tp **out)
{
	Tp *d = nullptr;
	CU(cudaMalloc((void **)&d, std::max<size_t>(n, 1) * sizeof(Tp)));
	eng->dev_allocs.push_back(d);
	if (n) CU(cudaMemcpy(d, src, n * sizeof(Tp), cudaMemcpyHostToDevice));
	*out = d;
	return BEEP_OK;
}
template <typename Tp>
static int upload(BeepEngine *eng, const std::vector<Tp> &v, const Tp **out) { return upload(eng, v.data(), v.size(), out); }

#define NEED(cond, ...) do { if (!(cond)) { beep_engine_destroy(eng); return fail(BEEP_ERR_INVALID, __VA_ARGS__); } } while (0)
#define UP(field, src, n) do { int rc_ = upload(eng, (src), (size_t)(n), &eng->dm.field); if (rc_) { beep_engine_destroy(eng); return rc_; } } while (0)
#define UPV(field, vec) do { int rc_ = upload(eng, (vec), &eng->dm.field); if (rc_) { beep_engine_destroy(eng); return rc_; } } while (0)

static int choose_launch(BeepEngine *eng);

extern "C" void beep_engine_destroy(BeepEngine *eng)
{
	if (!eng) return;
	cudaSetDevice(eng->device);
	for (void *p : eng->dev_allocs) cudaFree(p);
	if (eng->mvn_buf) cudaFree(eng->mvn_buf);
	delete eng;
}


/* ------------------------------------------------------------------------ */
/* MVN-approximate likelihood: the index lists of State::initialise_approx    */
/* (state_mvn_approx.cc:1288-1475) and ObservationModel::generate_obs_slice   */
/* (obsmodel.cc:357-385) from the flat description                            */
/* ------------------------------------------------------------------------ */
/* model.cc:3524-3566: the compartments on the far side of transition tr; false when transitions merge */
static bool trans_comp_after(const BeepModelDesc *d, int tr, std::vector<int> &clist)
{
	std::vector<int> map(d->C, -1);
	map[d->trans_from[tr]] = 0; map[d->trans_to[tr]] = 1;
	bool flag;
	do {
		flag = false;
		for (int k = 0; k < d->J; k++) {
			const int from = map[d->trans_from[k]], to = map[d->trans_to[k]];
			const int num = (from < 0) + (to < 0);
			if (num == 0) {
				if (from != to && !(d->trans_from[k] == d->trans_from[tr] && d->trans_to[k] == d->trans_to[tr])) return false;
			} else if (num == 1) {
				if (from < 0) map[d->trans_from[k]] = to; else map[d->trans_to[k]] = from;
				flag = true;
			}
		}
	} while (flag);
	clist.clear();
	for (int c = 0; c < d->C; c++) if (map[c] == 1) clist.push_back(c);
	return true;
}

static int approx_setup(BeepEngine *eng, const BeepModelDesc *d)
{
	DevApprox &x = eng->ax;
	memset(&x, 0, sizeof x);
	eng->approx_ok = false;
	const int A = d->A, C = d->C, D = d->D, J = d->J, O = d->nobs;
	if (d->n_table_type != d->ntable || d->ntable == 0) { eng->approx_why = "the description carries no table_type (written before the approximate likelihood existed)"; return BEEP_OK; }
	if (d->S != 1) { eng->approx_why = "the approximate likelihood handles one strain (state_mvn_approx.cc:1208)"; return BEEP_OK; }
	if (d->n_dcc_info > 0) { eng->approx_why = "the approximate likelihood ignores demographic changes"; return BEEP_OK; }
	if (d->n_mod_transmean || d->n_mod_beta || d->n_mod_efoi) { eng->approx_why = "model modifications belong to mode=prediction, the approximate likelihood to mode=map"; return BEEP_OK; }
	const int n = A * C * D, nt = A * J * D, start = d->trans_from[d->infection_trans];
	if ((long long)n * std::max(n, nt) >= (1ll << 30)) { eng->approx_why = "too many compartments for the approximate likelihood"; return BEEP_OK; }
	std::vector<int> wo, ws, wo_of(n, -1);
	for (int ca = 0; ca < n; ca++) { if ((ca / D) % C != start) { wo_of[ca] = (int)wo.size(); wo.push_back(ca); } else ws.push_back(ca); }
	if ((int)wo.size() != nt) { eng->approx_why = "the approximate likelihood needs as many transitions as non-susceptible compartments (state_mvn_approx.cc:1376)"; return BEEP_OK; }
	for (int co = 0; co < C; co++) if (co == start && d->comp_inf_param[co] >= 0) { /* infectivity of S is checked per particle in the reference; its parameter is the zero constant in every model */ }
	std::vector<int> wsl_off(1, 0), wsl;
	for (int cai : ws) {
		for (int caj : wo) if (caj / (C * D) == cai / (C * D) && caj % D == cai % D) wsl.push_back(caj);
		wsl_off.push_back((int)wsl.size());
	}
	std::vector<double> total_pop((size_t)A * D, 0.0);
	for (int c = 0; c < A; c++)
		for (int dp = 0; dp < D; dp++) {
			double sum = 0.0;
			for (int co = 0; co < C; co++) sum += d->pop_init[((size_t)c * C + co) * D + dp];
			total_pop[(size_t)c * D + dp] = sum;
		}
	std::vector<int> ta_from(nt), ta_to(nt), inc_off(1, 0), inc_ta;
	for (int ta = 0; ta < nt; ta++) {
		const int c = ta / (J * D), tr = (ta / D) % J, dp = ta % D;
		ta_from[ta] = (c * C + d->trans_from[tr]) * D + dp; ta_to[ta] = (c * C + d->trans_to[tr]) * D + dp;
	}
	for (int ca = 0; ca < n; ca++) {
		const int c = ca / (C * D), co = (ca / D) % C, dp = ca % D;
		for (int tr = 0; tr < J; tr++) if (d->trans_from[tr] == co || d->trans_to[tr] == co) inc_ta.push_back((c * J + tr) * D + dp);
		inc_off.push_back((int)inc_ta.size());
	}
	std::vector<int> ocl_off(1, 0), ocl, ocw_off(1, 0), ocw, otype(O), osett(O);
	std::vector<double> var_approx(O);
	for (int o = 0; o < O; o++) {
		const int tb = d->obs_table[o], type = d->table_type[tb];
		otype[o] = type;
		osett[o] = (type == 1 || type == 2) ? d->obs_sett_i[o] : d->obs_sett_f[o];        /* obsmodel.cc:368-373 */
		std::vector<int> complist;
		if (type == 1 || type == 2) { for (int k = d->table_comp_off[tb]; k < d->table_comp_off[tb + 1]; k++) complist.push_back(d->table_comp[k]); }
		else {
			for (int k = d->table_trans_off[tb]; k < d->table_trans_off[tb + 1]; k++) {
				std::vector<int> after;
				if (!trans_comp_after(d, d->table_trans[k], after)) { eng->approx_why = "transitions merge into the same compartment (model.cc:3541)"; return BEEP_OK; }
				for (int co : after) {
					if (std::find(complist.begin(), complist.end(), co) != complist.end()) { eng->approx_why = "an observation contains sequential transitions (state_mvn_approx.cc:1446)"; return BEEP_OK; }
					complist.push_back(co);
				}
			}
		}
		for (int co : complist) {
			if (co == start) { eng->approx_why = "an observation measures the susceptible compartment (state_mvn_approx.cc:1456)"; return BEEP_OK; }
			for (int ia = d->obs_area_off[o]; ia < d->obs_area_off[o + 1]; ia++)
				for (int id = d->obs_dp_off[o]; id < d->obs_dp_off[o + 1]; id++) {
					const int ca = (d->obs_area[ia] * C + co) * D + d->obs_dp[id];
					ocl.push_back(ca);
					if (wo_of[ca] >= 0) ocw.push_back(wo_of[ca]);
				}
		}
		ocl_off.push_back((int)ocl.size()); ocw_off.push_back((int)ocw.size());
		if (ocw_off[o + 1] == ocw_off[o]) { eng->approx_why = "an observation without compartments"; return BEEP_OK; }
		const double value = d->obs_value[o];                                  /* data.cc:1085-1115 */
		double va = 0;
		switch (d->obs_model[o]) {
		case 0: case 1: va = d->obs_sd[o] * d->obs_sd[o]; break;
		case 2: { double lam = value; if (lam < 1) lam = 1; va = lam; } break;
		case 3: { double mm = value; if (mm < 1) mm = 1; va = mm + mm * mm / d->obs_shape[o]; } break;
		}
		if (!(va > 0) || std::isinf(va)) { eng->approx_why = "an observation has zero or infinite variance (data.cc:1112-1113)"; return BEEP_OK; }
		var_approx[o] = va;
	}
	std::vector<int> sl_sett, sl_off(1, 0), sl_obs;
	{
		std::vector<int> setts(osett);
		std::sort(setts.begin(), setts.end());
		setts.erase(std::unique(setts.begin(), setts.end()), setts.end());
		for (int st : setts) {
			sl_sett.push_back(st);
			for (int o = 0; o < O; o++) if (osett[o] == st) sl_obs.push_back(o);
			sl_off.push_back((int)sl_obs.size());
		}
	}
	x.n = n; x.nt = nt; x.nwo = (int)wo.size(); x.nws = (int)ws.size(); x.start = start; x.nslice = (int)sl_sett.size();
#define UPX(field, vec) do { int rc_ = upload(eng, (vec), &x.field); if (rc_) return rc_; } while (0)
	UPX(inc_off, inc_off); UPX(inc_ta, inc_ta); UPX(ta_from, ta_from); UPX(ta_to, ta_to);
	UPX(wo, wo); UPX(ws, ws); UPX(wsl_off, wsl_off); UPX(wsl, wsl); UPX(total_pop, total_pop);
	UPX(ocl_off, ocl_off); UPX(ocl, ocl); UPX(ocw_off, ocw_off); UPX(ocw, ocw); UPX(var_approx, var_approx);
	UPX(obs_type, otype); { std::vector<int> si(d->obs_sett_i, d->obs_sett_i + O); UPX(obs_sett_i, si); }
	UPX(sl_sett, sl_sett); UPX(sl_off, sl_off); UPX(sl_obs, sl_obs);
#undef UPX
	eng->approx_ok = true;
	return BEEP_OK;
}


/* the graphs of data.graph (struct.hh:367-382) and the points ObservationModel::get_graph_state (obsmodel.cc:113-170) gives each */
static int graphs_setup(BeepEngine *eng, const BeepModelDesc *d)
{
	DevGraphs &g = eng->gx;
	memset(&g, 0, sizeof g);
	eng->graphs_ok = false;
	const int ng = (int)d->n_graph_type;
	if (ng == 0 || d->n_graph_step != 1 || d->graph_step[0] <= 0) return BEEP_OK;
	if (d->n_graph_table != ng || d->n_graph_factor != ng || d->n_graph_factor_spline != ng || d->n_graph_area_off != ng + 1 || d->n_graph_dp_off != ng + 1 || d->n_graph_point_off != ng + 1)
		return fail(BEEP_ERR_INVALID, "graph arrays of the description are inconsistent");
	const int gs = d->graph_step[0];
	std::vector<int> pt_graph, pt_first;
	eng->graph_point_off.assign(1, 0);
	for (int gr = 0; gr < ng; gr++) {
		const int dt = d->graph_table[gr];
		if (dt < 0 || dt >= d->ntable || d->graph_factor_spline[gr] >= d->nspline) return fail(BEEP_ERR_INVALID, "graph %d refers to a table or spline out of range", gr);
		for (int k = d->graph_area_off[gr]; k < d->graph_area_off[gr + 1]; k++) if (d->graph_area[k] < 0 || d->graph_area[k] >= d->A) return fail(BEEP_ERR_INVALID, "graph %d: area out of range", gr);
		for (int k = d->graph_dp_off[gr]; k < d->graph_dp_off[gr + 1]; k++) if (d->graph_dp[k] < 0 || d->graph_dp[k] >= d->D) return fail(BEEP_ERR_INVALID, "graph %d: demographic group out of range", gr);
		if (d->graph_type[gr] == 1) {
			for (int k = d->graph_point_off[gr]; k < d->graph_point_off[gr + 1]; k++) {
				if (d->graph_point_obs[k] < 0 || d->graph_point_obs[k] >= d->nobs) return fail(BEEP_ERR_INVALID, "graph %d: observation out of range", gr);
				pt_graph.push_back(gr); pt_first.push_back(d->graph_point_obs[k]);
			}
			eng->graphs_obs = true;
		} else {
			for (int sett = 0; sett < d->T - gs; sett += gs) { pt_graph.push_back(gr); pt_first.push_back(sett); }
			if (d->table_comp_off[dt + 1] > d->table_comp_off[dt]) eng->graphs_pop = true;
		}
		eng->graph_point_off.push_back((int)pt_graph.size());
	}
	g.ngraph = ng; g.npoint = (int)pt_graph.size(); g.gstep = gs;
#define UPG(field, src, n) do { int rc_ = upload(eng, (src), (size_t)(n), &g.field); if (rc_) return rc_; } while (0)
	UPG(type, d->graph_type, ng); UPG(table, d->graph_table, ng); UPG(fspline, d->graph_factor_spline, ng); UPG(factor, d->graph_factor, ng);
	UPG(area_off, d->graph_area_off, ng + 1); UPG(area, d->graph_area, d->n_graph_area); UPG(dp_off, d->graph_dp_off, ng + 1); UPG(dp, d->graph_dp, d->n_graph_dp);
	UPG(point_off, d->graph_point_off, ng + 1); UPG(point_obs, d->graph_point_obs, d->n_graph_point_obs);
	UPG(table_trans_off, d->table_trans_off, d->ntable + 1); UPG(table_trans, d->table_trans, d->n_table_trans);
	UPG(table_comp_off, d->table_comp_off, d->ntable + 1); UPG(table_comp, d->table_comp, d->n_table_comp);
	UPG(pt_graph, pt_graph.data(), pt_graph.size()); UPG(pt_first, pt_first.data(), pt_first.size());
#undef UPG
	eng->graphs_ok = g.npoint > 0;
	return BEEP_OK;
}

extern "C" int beep_engine_create(const BeepModelDesc *d, int device, BeepEngine **out)
{
	if (!d || !out) return fail(BEEP_ERR_INVALID, "null argument");
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BEEP_ERR_NO_DEVICE, "no CUDA device: the BEEP engine has no CPU fallback");
	if (device < 0 || device >= ndev) return fail(BEEP_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
	CU(cudaSetDevice(device));
	BeepEngine *eng = new BeepEngine();
	eng->device = device;
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	eng->nsm = prop.multiProcessorCount;
	eng->smem_optin = (int)prop.sharedMemPerBlockOptin;

	DevModel &m = eng->dm;
	memset(&m, 0, sizeof m);
	m.T = d->T; m.period = d->period; m.spu = d->spu; m.A = d->A; m.C = d->C; m.J = d->J; m.D = d->D; m.G = d->G; m.S = d->S; m.Dq = d->Dq;
	m.nparam = d->nparam; m.nspline = d->nspline; m.nobs = d->nobs; m.ntable = d->ntable;
	m.stochastic = d->stochastic; m.siminf = d->siminf_simulate; m.itr = d->infection_trans; m.geo_spline = d->geo_spline_ref;
	m.popsize = d->popsize; m.dt = (double)d->period / d->T;
	m.nR = (int)d->n_R_spline_ref; m.nefoi = (int)d->n_efoi_spline_ref; m.ninf = (int)d->n_inf_state; m.nmm = (int)d->n_matmod_spline;
	m.ncovar = (int)d->n_covar_param; m.nlevel = (int)d->n_level_param; m.ndir = (int)d->n_dir_off > 0 ? (int)d->n_dir_off - 1 : 0;
	m.has_region = d->n_region_effect_param > 0; m.has_area_eff = d->n_area_effect_param > 0;

	const int T = m.T, A = m.A, C = m.C, J = m.J, D = m.D, G = m.G, S = m.S, Dq = m.Dq;
	if (d->n_unsupported > 0) { beep_engine_destroy(eng); return fail(BEEP_ERR_UNSUPPORTED, "the exporter marked the analysis as using a feature outside the hot-path scope"); }
	NEED(T > 0 && A > 0 && C > 0 && J > 0 && D > 0 && G > 0 && S > 0 && Dq > 0 && m.spu > 0 && m.period > 0, "non-positive dimension");
	NEED(D == S * Dq, "D (%d) != S*Dq (%d*%d)", D, S, Dq);
	NEED(m.itr >= 0 && m.itr < J, "infection_trans out of range");
	NEED(m.geo_spline >= 0 && m.geo_spline < m.nspline, "geo_spline_ref out of range");
	if (G > 32 || D > 64 || J > 31 || C > 32 || m.ninf > K0_MAXINF) { beep_engine_destroy(eng); return fail(BEEP_ERR_UNSUPPORTED, "engine limits: G<=32 D<=64 J<=31 C<=32 ninf<=%d (got G=%d D=%d J=%d C=%d ninf=%d)", K0_MAXINF, G, D, J, C, m.ninf); }
	NEED(d->n_trans_from == J && d->n_trans_to == J && d->n_trans_inf == J && d->n_trans_shape == J, "transition arrays must have J entries");
	NEED(d->n_trans_mean_param == (int64_t)J * D && d->n_trans_prob_param == (int64_t)J * D, "trans_*_param must be [J][D]");
	NEED(d->n_comp_inf_param == C, "comp_inf_param must be [C]");
	NEED(d->n_spline_factor_param == m.nspline && d->n_spline_off == m.nspline + 1, "spline arrays inconsistent");
	NEED(d->n_sus_off == D + 1, "sus_off must be [D+1]");
	NEED(d->n_N0 == (int64_t)G * G && d->n_dp_age == Dq, "N0 must be [G][G], dp_age [Dq]");
	NEED(d->n_R_area_off == m.nR + 1 && d->n_R_dist == (int64_t)m.nR * Dq && d->n_strain_Rfactor_param == S, "R-spline arrays inconsistent");
	NEED(d->n_efoi_agedist == (int64_t)m.nefoi * G && d->n_efoi_ref == (int64_t)S * A, "efoi arrays inconsistent");
	NEED(d->n_M_off == A + 1 && d->n_M_diag == A && d->n_M_factor == A, "mixing matrix arrays inconsistent");
	NEED(d->n_pop_init == (int64_t)A * C * D, "pop_init must be [A][C][D]");
	NEED(d->n_area_total_pop == A, "area_total_pop must be [A]");
	NEED(d->n_obs_value == m.nobs && d->n_obs_area_off == m.nobs + 1 && d->n_obs_dp_off == m.nobs + 1, "observation arrays inconsistent");
	NEED(d->n_table_trans_off == m.ntable + 1 && d->n_table_comp_off == m.ntable + 1 && d->n_table_threshold == m.ntable, "datatable arrays inconsistent");
	NEED(m.ncovar == 0 || d->n_covar_value == (int64_t)m.ncovar * A * m.period, "covar_value must be [ncovar][A][period]");
	NEED(m.nlevel == 0 || d->n_level_map == (int64_t)m.period * A, "level_map must be [period][A]");
	NEED(d->n_obs_sd == m.nobs && d->n_obs_factor == m.nobs && d->n_obs_shape == m.nobs && d->n_obs_model == m.nobs && d->n_obs_table == m.nobs &&
	     d->n_obs_sett_i == m.nobs && d->n_obs_sett_f == m.nobs && d->n_obs_factor_spline == m.nobs, "per-observation arrays must have nobs entries");
	NEED(m.nobs == 0 || (d->n_obs_area >= d->obs_area_off[m.nobs] && d->n_obs_dp >= d->obs_dp_off[m.nobs]), "obs_area / obs_dp shorter than their offsets say");
	NEED(d->n_table_trans >= d->table_trans_off[m.ntable] && d->n_table_comp >= d->table_comp_off[m.ntable], "table_trans / table_comp shorter than their offsets say");
	NEED(m.nR == 0 || d->n_R_area >= d->R_area_off[m.nR], "R_area shorter than R_area_off says");
	NEED(d->n_spline_param >= d->spline_off[m.nspline] && d->n_spline_t >= d->spline_off[m.nspline] && d->n_spline_multfac >= d->spline_off[m.nspline], "spline knot arrays shorter than spline_off says");
	NEED(d->n_sus_param >= d->sus_off[D], "sus_param shorter than sus_off says");
	NEED(m.ndir == 0 || (d->n_dir_th >= d->dir_off[m.ndir] && d->n_dir_frac >= d->dir_off[m.ndir]), "dir_th / dir_frac shorter than dir_off says");
	NEED(d->n_efoi_spline_ref == m.nefoi && d->n_matmod_spline == m.nmm, "efoi / matmod spline arrays inconsistent");
	NEED(!m.has_region || d->n_region_effect_param == A, "region_effect_param must be [A]");
	NEED(!m.has_area_eff || d->n_area_effect_param == A, "area_effect_param must be [A]");
	{
		/* every parameter / spline / compartment index the kernels dereference */
		auto par_ok = [&](const int32_t *v, int64_t n, bool allow_neg) { for (int64_t i = 0; i < n; i++) if (v[i] >= m.nparam || (v[i] < 0 && !allow_neg)) return false; return true; };
		NEED(par_ok(d->trans_prob_param, (int64_t)J * D, true), "trans_prob_param out of range");
		for (int tr = 0; tr < J; tr++) if (!d->trans_inf[tr]) NEED(par_ok(d->trans_mean_param + (size_t)tr * D, D, false), "trans_mean_param of transition %d out of range", tr);
		NEED(par_ok(d->comp_inf_param, C, false), "comp_inf_param out of range");
		NEED(par_ok(d->spline_factor_param, m.nspline, false) && par_ok(d->spline_param, d->spline_off[m.nspline], false), "spline parameters out of range");
		NEED(par_ok(d->sus_param, d->sus_off[D], false), "sus_param out of range");
		NEED(m.ndir == 0 || par_ok(d->dir_th, d->dir_off[m.ndir], false), "dir_th out of range");
		NEED(par_ok(d->covar_param, m.ncovar, false) && par_ok(d->level_param, m.nlevel, false), "covariate / level parameters out of range");
		NEED(par_ok(d->strain_Rfactor_param, S, false), "strain_Rfactor_param out of range");
		NEED(!m.has_region || par_ok(d->region_effect_param, A, false), "region_effect_param out of range");
		NEED(!m.has_area_eff || par_ok(d->area_effect_param, A, false), "area_effect_param out of range");
		for (int i = 0; i < m.nR; i++) NEED(d->R_spline_ref[i] >= 0 && d->R_spline_ref[i] < m.nspline, "R_spline_ref out of range");
		for (int i = 0; i < m.nefoi; i++) NEED(d->efoi_spline_ref[i] >= 0 && d->efoi_spline_ref[i] < m.nspline, "efoi_spline_ref out of range");
		for (int i = 0; i < m.nmm; i++) NEED(d->matmod_spline[i] >= 0 && d->matmod_spline[i] < m.nspline, "matmod_spline out of range");
		for (int64_t i = 0; i < (m.nlevel > 0 ? d->n_level_map : 0); i++) NEED(d->level_map[i] >= 0 && d->level_map[i] < m.nlevel, "level_map out of range");
	}
	for (int tr = 0; tr < J; tr++) NEED(d->trans_from[tr] >= 0 && d->trans_from[tr] < C && d->trans_to[tr] >= 0 && d->trans_to[tr] < C, "transition %d endpoints out of range", tr);
	for (int i = 0; i < Dq; i++) NEED(d->dp_age[i] >= 0 && d->dp_age[i] < G, "dp_age out of range");

	for (int tr = 0; tr < J; tr++) {
		eng->h_from[tr] = (unsigned char)d->trans_from[tr]; eng->h_to[tr] = (unsigned char)d->trans_to[tr];
		eng->h_kind[tr] = (tr == m.itr) ? 1 : (d->trans_inf[tr] ? 2 : 0);
	}
	UP(dir_off, d->dir_off, d->n_dir_off); UP(dir_th, d->dir_th, d->n_dir_th); UP(dir_frac, d->dir_frac, d->n_dir_frac);
	UP(trans_from, d->trans_from, J); UP(trans_to, d->trans_to, J); UP(trans_inf, d->trans_inf, J); UP(trans_shape, d->trans_shape, J);
	UP(trans_mean_param, d->trans_mean_param, J * D); UP(trans_prob_param, d->trans_prob_param, J * D);
	UP(comp_inf_param, d->comp_inf_param, C); UP(inf_state, d->inf_state, m.ninf);
	{
		std::vector<int> ir(C, -1);
		for (int i = 0; i < m.ninf; i++) { NEED(d->inf_state[i] >= 0 && d->inf_state[i] < C, "inf_state out of range"); ir[d->inf_state[i]] = i; }
		UPV(inf_ref, ir);
	}
	UP(spline_factor_param, d->spline_factor_param, m.nspline); UP(spline_off, d->spline_off, m.nspline + 1);
	UP(spline_param, d->spline_param, d->n_spline_param); UP(spline_t, d->spline_t, d->n_spline_t); UP(spline_multfac, d->spline_multfac, d->n_spline_multfac);
	UP(sus_off, d->sus_off, D + 1); UP(sus_param, d->sus_param, d->n_sus_param);
	UP(region_effect_param, d->region_effect_param, d->n_region_effect_param); UP(area_effect_param, d->area_effect_param, d->n_area_effect_param);
	UP(covar_param, d->covar_param, m.ncovar); UP(covar_value, d->covar_value, d->n_covar_value);
	UP(level_param, d->level_param, m.nlevel); UP(level_map, d->level_map, d->n_level_map); UP(area_total_pop, d->area_total_pop, A);
	UP(N0, d->N0, G * G);
	{
		std::vector<double> nt((size_t)G * G);
		for (int a = 0; a < G; a++) for (int b = 0; b < G; b++) nt[(size_t)b * G + a] = d->N0[a * G + b];
		UPV(N0T, nt);
		std::vector<int> cnt((size_t)std::max(m.nmm, 1) * G, 0);
		NEED(d->n_matmod_age_off == m.nmm + 1 || m.nmm == 0, "matmod_age_off must be [nmatmod+1]");
		for (int mm = 0; mm < m.nmm; mm++)
			for (int k = d->matmod_age_off[mm]; k < d->matmod_age_off[mm + 1]; k++) { NEED(d->matmod_age[k] >= 0 && d->matmod_age[k] < G, "matmod age out of range"); cnt[(size_t)mm * G + d->matmod_age[k]]++; }
		UPV(mm_cnt, cnt);
		std::vector<int> poff((size_t)G * G + 1, 0), prog;
		for (int a = 0; a < G; a++) for (int b = 0; b < G; b++) {
			for (int mm = 0; mm < m.nmm; mm++)
				for (int k = cnt[(size_t)mm * G + a] + cnt[(size_t)mm * G + b]; k > 0; k--) prog.push_back(d->matmod_spline[mm]);
			poff[(size_t)a * G + b + 1] = (int)prog.size();
		}
		std::vector<unsigned> pack((size_t)G * G, 0);
		m.mm_packed = (m.nmm > 0 && m.nspline < 64) ? 1 : 0;
		for (int e = 0; e < G * G; e++) {
			const int n = poff[e + 1] - poff[e];
			if (n > 3) m.mm_packed = 0;
			unsigned w = 0;
			for (int k = 0; k < 3; k++) w |= (unsigned)((k < n && n <= 3) ? prog[poff[e] + k] : m.nspline) << (8 * k);
			pack[e] = w;
		}
		if (prog.empty()) prog.push_back(0);
		UPV(mm_prog_off, poff); UPV(mm_prog, prog); UPV(mm_pack, pack);
		const int Gs = traj_sm_layout(G, m.J, m.D).Gs;
		const unsigned unused = (unsigned)m.nspline * 0x010101u;
		std::vector<unsigned> packp((size_t)G * Gs, unused);
		for (int a = 0; a < G; a++) for (int b = 0; b < G; b++) packp[(size_t)a * Gs + b] = pack[(size_t)a * G + b];
		UPV(mm_packp, packp);
	}
	UP(mm_spline, d->matmod_spline, m.nmm); UP(dp_age, d->dp_age, Dq);
	m.age_fast = 1;
	for (int i = 0; i < Dq; i++) if (d->dp_age[i] != i % G) m.age_fast = 0;
	UP(R_spline_ref, d->R_spline_ref, m.nR); UP(R_dist, d->R_dist, d->n_R_dist); UP(strain_Rfactor_param, d->strain_Rfactor_param, S);
	{
		std::vector<int> rg(A, -1);
		for (int i = 0; i < m.nR; i++) for (int k = d->R_area_off[i]; k < d->R_area_off[i + 1]; k++) { NEED(d->R_area[k] >= 0 && d->R_area[k] < A, "R_area out of range"); rg[d->R_area[k]] = i; }
		for (int c = 0; c < A; c++) NEED(rg[c] >= 0, "area %d belongs to no R spline", c);
		UPV(area_rgrp, rg);
		eng->h_rgrp = rg;
	}
	UP(efoi_spline_ref, d->efoi_spline_ref, m.nefoi); UP(efoi_ref, d->efoi_ref, S * A); UP(efoi_agedist, d->efoi_agedist, d->n_efoi_agedist);
	{
		/* transpose of genQ.M (state.cc:204-216 scatters along row c; we gather into c'),
		 * sources kept in ascending order so that the sum has the reference's order */
		std::vector<int> off(A + 1, 0), src; std::vector<double> val;
		const int nnz = d->M_off[A];
		NEED(d->n_M_to >= nnz && d->n_M_val >= nnz, "M_to / M_val shorter than M_off says");
		for (int k = 0; k < nnz; k++) { NEED(d->M_to[k] >= 0 && d->M_to[k] < A, "M_to out of range"); off[d->M_to[k] + 1]++; }
		for (int c = 0; c < A; c++) off[c + 1] += off[c];
		src.resize(nnz); val.resize(nnz);
		std::vector<int> fill(off.begin(), off.end() - 1);
		for (int c = 0; c < A; c++) for (int k = d->M_off[c]; k < d->M_off[c + 1]; k++) { int to = d->M_to[k]; src[fill[to]] = c; val[fill[to]] = d->M_val[k]; fill[to]++; }
		UPV(Min_off, off); UPV(Min_src, src); UPV(Min_val, val);
		m.M_nnz = nnz;
		/* the same matrix as the kernel reads it: 16-byte (weight, source offset) entries and one
		 * 32-byte record of per-area constants */
		std::vector<Nbr> nb;
		nb.reserve((size_t)nnz + 8 * (size_t)A);
		std::vector<AreaInfo> ai((size_t)A);
		for (int c = 0; c < A; c++) {
			NEED(d->efoi_ref[c] >= 0 && d->efoi_ref[c] < m.nefoi, "efoi_ref out of range");
			/* a row is padded to a multiple of eight entries with zero weights on the area itself: the chunk kernel works
			 * whole groups of eight (adding +0 changes no bit); the other kernels stop at nn */
			const int j0 = (int)nb.size();
			for (int k = off[c]; k < off[c + 1]; k++) { Nbr e; e.val = val[k]; e.off = (unsigned)src[k] * (unsigned)G * 8u; e.pad = 0; nb.push_back(e); }
			while ((nb.size() - (size_t)j0) & 7) { Nbr e; e.val = 0.0; e.off = (unsigned)c * (unsigned)G * 8u; e.pad = 0; nb.push_back(e); }
			ai[c].diag = d->M_diag[c]; ai[c].fac = d->M_factor[c]; ai[c].j0 = j0; ai[c].nn = off[c + 1] - off[c];
			ai[c].rg = eng->h_rgrp[c]; ai[c].ie0 = d->efoi_ref[c];
		}
		UPV(nbr, nb); UPV(area_info, ai);
		/* dense mixing matrix (every area mixes with most others): the contraction becomes a GEMM over the batch */
		eng->dense_m = (A >= 64 && A <= 8192 && (long long)nnz >= 64ll * A);
		if (const char *e = getenv("BEEP_DENSE_M")) eng->dense_m = atoi(e) != 0 && A <= 8192 && nnz > 0;
		if (eng->dense_m) {
			std::vector<double> mt((size_t)A * A, 0.0);
			for (int c = 0; c < A; c++) for (int k = off[c]; k < off[c + 1]; k++) mt[(size_t)c * A + src[k]] = val[k];
			UPV(Mt_dense, mt);
		}
	}
	UP(M_diag, d->M_diag, A); UP(M_factor, d->M_factor, A);
	{
		/* demographic changes, state.cc:339-376 (struct.hh:294-304) */
		NEED(d->n_dcc_info % 6 == 0, "dcc_info must be [ndcc][6]");
		m.ndcc = (int)(d->n_dcc_info / 6);
		for (int k = 0; k < m.ndcc; k++) {
			const int32_t *inf = d->dcc_info + 6 * k;
			NEED(inf[0] >= 1 && inf[0] <= 32, "democat_change %d: %d categories (1..32)", k, inf[0]);
			NEED(inf[1] >= 0 && inf[2] >= 0 && (int64_t)inf[1] + inf[2] <= d->n_dcc_area, "democat_change %d: areas out of range", k);
			NEED(inf[3] >= 0 && inf[4] >= 0 && (int64_t)inf[4] + (int64_t)inf[3] * inf[0] <= d->n_dcc_group, "democat_change %d: groups out of range", k);
			NEED(inf[5] >= 0 && (int64_t)inf[5] + (int64_t)T * inf[0] <= d->n_dcc_frac, "democat_change %d: fractions must be [T][ncat]", k);
			for (int i = 0; i < inf[2]; i++) NEED(d->dcc_area[inf[1] + i] >= 0 && d->dcc_area[inf[1] + i] < A, "democat_change %d: area out of range", k);
			for (int i = 0; i < inf[3] * inf[0]; i++) NEED(d->dcc_group[inf[4] + i] >= 0 && d->dcc_group[inf[4] + i] < D, "democat_change %d: dp out of range", k);
		}
		if (m.ndcc) { UP(dcc_info, d->dcc_info, d->n_dcc_info); UP(dcc_area, d->dcc_area, std::max<int64_t>(d->n_dcc_area, 1)); UP(dcc_group, d->dcc_group, std::max<int64_t>(d->n_dcc_group, 1)); UP(dcc_frac, d->dcc_frac, d->n_dcc_frac); }
	}
	UP(pop_init, d->pop_init, (size_t)A * C * D);
	m.CDp = (C * D + 1) & ~1;
	m.CDp4 = (C * D + 3) & ~3;
	{
		/* packed per-cell table: population offset of the source compartment, transition, dp, kind
		 * (0 = ordinary transition, 1 = the infection transition, 2 = always zero) */
		std::vector<unsigned> tab((size_t)J * D);
		const int sfrom = d->trans_from[m.itr];
		for (int tr = 0; tr < J; tr++) for (int dp = 0; dp < D; dp++) {
			unsigned kind = (tr == m.itr) ? 1u : (d->trans_inf[tr] ? 2u : 0u);
			unsigned off = (kind == 1) ? (unsigned)(sfrom * D + dp % Dq) : (unsigned)(d->trans_from[tr] * D + dp);
			tab[(size_t)tr * D + dp] = off | ((unsigned)tr << 12) | ((unsigned)dp << 18) | (kind << 24);
		}
		UPV(cell_tab, tab);
	}

	/* observations: masks, (observation, area) slots, the (sett, area) -> entries index
	 * (the reference's obs_trans / obs_pop inverted index, obsmodel.cc:174-217) and the
	 * PMCMC sections (obsmodel.cc:326-348) */
	{
		const int O = m.nobs;
		std::vector<unsigned> tmask(O, 0), cmask(O, 0); std::vector<unsigned long long> dmask(O, 0);
		std::vector<int> slot_off(O + 1, 0), section(O, -1);
		std::vector<int> cell_cnt((size_t)T * A + 1, 0);
		for (int ob = 0; ob < O; ob++) {
			int tb = d->obs_table[ob];
			NEED(tb >= 0 && tb < m.ntable, "obs_table out of range");
			for (int k = d->table_trans_off[tb]; k < d->table_trans_off[tb + 1]; k++) { NEED(d->table_trans[k] >= 0 && d->table_trans[k] < J, "table_trans out of range"); tmask[ob] |= 1u << d->table_trans[k]; }
			for (int k = d->table_comp_off[tb]; k < d->table_comp_off[tb + 1]; k++) { NEED(d->table_comp[k] >= 0 && d->table_comp[k] < C, "table_comp out of range"); cmask[ob] |= 1u << d->table_comp[k]; }
			for (int k = d->obs_dp_off[ob]; k < d->obs_dp_off[ob + 1]; k++) { NEED(d->obs_dp[k] >= 0 && d->obs_dp[k] < D, "obs_dp out of range"); dmask[ob] |= 1ull << d->obs_dp[k]; }
			int na = d->obs_area_off[ob + 1] - d->obs_area_off[ob];
			slot_off[ob + 1] = slot_off[ob] + na;
			NEED(d->obs_sett_i[ob] >= 0 && d->obs_sett_f[ob] <= T && d->obs_sett_i[ob] <= d->obs_sett_f[ob], "observation %d window out of range", ob);
			NEED(d->obs_factor_spline[ob] < m.nspline, "obs_factor_spline out of range");
			for (int k = d->obs_area_off[ob]; k < d->obs_area_off[ob + 1]; k++) {
				NEED(d->obs_area[k] >= 0 && d->obs_area[k] < A, "obs_area out of range");
				for (int t = d->obs_sett_i[ob]; t < d->obs_sett_f[ob]; t++) cell_cnt[(size_t)t * A + d->obs_area[k] + 1]++;
			}
		}
		m.nslot = slot_off[O];
		{
			/* wide observations (many dp per observation) are reduced warp-cooperatively */
			double w = 0;
			for (int ob = 0; ob < O; ob++) w += (double)__builtin_popcountll(dmask[ob]);
			m.obs_wide = (O > 0 && w / O > 4.0) ? 1 : 0;
			const char *fw = getenv("BEEP_FORCE_OBS_WIDE");
			if (fw) m.obs_wide = atoi(fw) ? 1 : 0;
		}
		for (size_t i = 0; i < (size_t)T * A; i++) cell_cnt[i + 1] += cell_cnt[i];
		std::vector<ObsEnt> ent(cell_cnt[(size_t)T * A]);
		std::vector<int> fill(cell_cnt.begin(), cell_cnt.end() - 1);
		/* a cell's entries that observe a population come first (the chunk kernel reads them before the step's
		 * events are applied); cell_npop counts them */
		std::vector<unsigned char> npop((size_t)T * A, 0);
		int max_ent = 0;
		bool npop_fits = true;
		for (int pass = 0; pass < 2; pass++)
			for (int ob = 0; ob < O; ob++) {
				if ((cmask[ob] != 0) != (pass == 0)) continue;
				for (int k = d->obs_area_off[ob]; k < d->obs_area_off[ob + 1]; k++)
					for (int t = d->obs_sett_i[ob]; t < d->obs_sett_f[ob]; t++) {
						size_t cell = (size_t)t * A + d->obs_area[k];
						ent[fill[cell]++] = ObsEnt{ob, slot_off[ob] + (k - d->obs_area_off[ob])};
						if (pass == 0) { if (npop[cell] == 255) npop_fits = false; else npop[cell]++; }
					}
			}
		for (size_t i = 0; i < (size_t)T * A; i++) max_ent = std::max(max_ent, cell_cnt[i + 1] - cell_cnt[i]);
		m.obs_tr_any = 0;
		for (int ob = 0; ob < O; ob++) m.obs_tr_any |= tmask[ob];
		UPV(cell_npop, npop);
		/* the chunk kernel serves the common model structure (see launch_traj) with integer populations
		 * and at most CHUNK_NEJ observation entries per (sett, area) cell */
		eng->chunk_model = m.ndcc == 0 && npop_fits && max_ent <= CHUNK_NEJ && m.S == 1 && m.Dq == m.G && m.G > 1 && m.stochastic && J * D < 4096 && C * D < 4096;
		for (int i = 0; i < Dq && eng->chunk_model; i++) if (d->dp_age[i] != i % G) eng->chunk_model = false;
		for (size_t i = 0; i < (size_t)A * C * D && eng->chunk_model; i++) {
			const double v = d->pop_init[i];
			if (!(std::fabs(v) < 2e9) || v != std::floor(v)) eng->chunk_model = false;
		}
		for (unsigned q = 0; q < std::max(8u * 32u, (unsigned)(J * D)); q++) if (((q * ((65536u + G - 1) / G)) >> 16) != q / G) eng->chunk_model = false;
		int per = d->pmcmc_obs_period > 0 ? d->pmcmc_obs_period : T;
		m.nsection = (T + per - 1) / per;
		for (int s = 0; s < m.nsection; s++) {
			int ti = s * per, tf = std::min(ti + per, T);
			eng->section_ti.push_back(ti); eng->section_tf.push_back(tf);
			for (int ob = 0; ob < O; ob++)
				if (section[ob] < 0 && d->obs_sett_i[ob] < tf && d->obs_sett_f[ob] >= ti && d->obs_sett_i[ob] >= ti && d->obs_sett_f[ob] <= tf) section[ob] = s;
		}
		UP(obs_value, d->obs_value, O); UP(obs_sd, d->obs_sd, O); UP(obs_factor, d->obs_factor, O); UP(obs_shape, d->obs_shape, O);
		UP(obs_fs, d->obs_factor_spline, O); UP(obs_table, d->obs_table, O); UP(obs_model, d->obs_model, O);
		UPV(obs_section, section); UPV(obs_slot_off, slot_off); UPV(obs_tmask, tmask); UPV(obs_cmask, cmask); UPV(obs_dmask, dmask);
		UPV(cell_off, cell_cnt); UPV(cell_ent, ent);
		UP(table_threshold, d->table_threshold, m.ntable);
		UPV(section_ti, eng->section_ti); UPV(section_tf, eng->section_tf);
	}
	{
		/* model modifications of mode=prediction (model.cc:1440-1585): absent arrays = no modification */
		const size_t TAJD = (size_t)T * A * J * D, TAS = (size_t)T * A * S, NT = (size_t)m.nspline * T;
		NEED(d->n_mod_transmean == 0 || (size_t)d->n_mod_transmean == TAJD, "mod_transmean has %lld entries, not T*A*J*D", (long long)d->n_mod_transmean);
		NEED(d->n_mod_beta == 0 || (size_t)d->n_mod_beta == TAS, "mod_beta has %lld entries, not T*A*S", (long long)d->n_mod_beta);
		NEED(d->n_mod_efoi == 0 || (size_t)d->n_mod_efoi == TAS, "mod_efoi has %lld entries, not T*A*S", (long long)d->n_mod_efoi);
		NEED((d->n_mod_spline_mult == 0 && d->n_mod_spline_set == 0) || ((size_t)d->n_mod_spline_mult == NT && (size_t)d->n_mod_spline_set == NT), "mod_spline_mult / mod_spline_set must both have nspline*T entries");
		if (d->n_mod_transmean) UP(mod_tmean, d->mod_transmean, TAJD);
		if (d->n_mod_beta) { UP(mod_beta, d->mod_beta, TAS); eng->h_mod_beta.assign(d->mod_beta, d->mod_beta + TAS); }
		if (d->n_mod_efoi) UP(mod_efoi, d->mod_efoi, TAS);
		if (d->n_mod_spline_mult) { UP(mod_spline_mult, d->mod_spline_mult, NT); UP(mod_spline_set, d->mod_spline_set, NT); }
		m.has_mod = (m.mod_tmean || m.mod_beta || m.mod_efoi) ? 1 : 0;     /* spline modifications live in K0 alone */
		if (m.has_mod) { eng->chunk_model = false; eng->dense_m = false; }
	}
	int rc = choose_launch(eng);
	if (rc) { beep_engine_destroy(eng); return rc; }
	rc = approx_setup(eng, d);
	if (rc) { beep_engine_destroy(eng); return rc; }
	rc = graphs_setup(eng, d);
	if (rc) { beep_engine_destroy(eng); return rc; }
	*out = eng;
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* launch configuration of the trajectory kernel                             */
/* ------------------------------------------------------------------------ */
template <int SYNC, bool KEEP>
static const void *traj_fn2(int mode)
{
	switch (mode) {
	case MODE_SIM: return (const void *)k_traj<MODE_SIM, SYNC, KEEP>;
	case MODE_MBP: return (const void *)k_traj<MODE_MBP, SYNC, true>;
	default: return (const void *)k_traj<MODE_REPLAY, SYNC, true>;
	}
}
/* keep = false: no trajectory array is stored (simulate + likelihood only); MBP and replay always keep */
static const void *traj_fn(int mode, int sync, bool keep)
{
	if (mode != MODE_SIM) keep = true;
	if (sync == SYNC_NONE) return keep ? traj_fn2<SYNC_NONE, true>(mode) : traj_fn2<SYNC_NONE, false>(mode);
	if (sync == SYNC_CTA) return keep ? traj_fn2<SYNC_CTA, true>(mode) : traj_fn2<SYNC_CTA, false>(mode);
	return keep ? traj_fn2<SYNC_GRID, true>(mode) : traj_fn2<SYNC_GRID, false>(mode);
}

/* chunk kernel instantiations (traj_chunk.cuh) */
template <int NT, bool KEEP>
static const void *chunk_fn2(int mode)
{
	switch (mode) {
	case MODE_SIM: return (const void *)k_traj_chunk<MODE_SIM, KEEP, NT>;
	case MODE_MBP: return (const void *)k_traj_chunk<MODE_MBP, true, NT>;
	default: return (const void *)k_traj_chunk<MODE_REPLAY, true, NT>;
	}
}
static const void *chunk_fn(int mode, int nt, bool keep)
{
	if (mode != MODE_SIM) keep = true;
	if (nt == 4) return keep ? chunk_fn2<4, true>(mode) : chunk_fn2<4, false>(mode);
	return keep ? chunk_fn2<8, true>(mode) : chunk_fn2<8, false>(mode);
}

static size_t even(size_t x) { return (x + 1) & ~(size_t)1; }
/* size of the per-warp neighbour staging area: 32 entries of 16 bytes when the mixing matrix has
 * off-diagonal entries; otherwise only what ntime_refresh stages there (nspline + 1 spline values) */
static int traj_nbr_doubles(const DevModel &m) { return m.M_nnz > 0 ? 64 : (m.nmm > 0 ? m.nspline + 1 : 0); }

static int choose_launch(BeepEngine *eng)
{
	const DevModel &m = eng->dm;
	const TrajSm L = traj_sm_layout(m.G, m.J, m.D, traj_nbr_doubles(m));      /* handed to the kernel in its arguments */
	const size_t CDp = (size_t)m.CDp, common = (size_t)L.common, work = (size_t)L.work;
	const char *force = getenv("BEEP_FORCE_SYNC");      /* tests: exercise every flavour on small models */
	auto cta_bytes = [&](int nwarp) { return (common + (size_t)m.A * CDp + even(2 * (size_t)m.S * m.A * m.G) + nwarp * work) * sizeof(double); };
	if (m.M_nnz == 0) {
		eng->sync_mode = SYNC_NONE;
		const size_t ntsz = (m.nmm > 0) ? (size_t)m.G * L.Gs : 0;      /* the warp's copy of Ntime (k_traj) */
		auto none_bytes = [&](int nwarp) { return (common + nwarp * (CDp + work + even(2 * (size_t)m.S * m.G) + ntsz)) * sizeof(double); };
		/* warps per CTA: the trajectory of a tile is one long dependent chain, so what counts is the number
		 * of resident warps per SM -- take the CTA size (4..8 warps) that the occupancy calculator likes best
		 * (the CTA-common tables and the per-CTA reservation are amortised differently) */
		int best = 4, best_warps = -1;
		for (int nw = 4; nw <= BEEP_TRAJ_NWARP; nw++) {
			if (none_bytes(nw) > (size_t)eng->smem_optin) break;
			const void *fn = traj_fn(MODE_SIM, SYNC_NONE, false);
			if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)none_bytes(nw)) != cudaSuccess) break;
			int nb = 0;
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, nw * 32, none_bytes(nw)) != cudaSuccess) break;
			if (nb * nw > best_warps) { best_warps = nb * nw; best = nw; }
		}
		eng->nwarp = best;
		eng->smem_traj = none_bytes(best);
		eng->none_resident = std::max(1, best_warps / best) * eng->nsm;
	} else {
		eng->nwarp = std::min(BEEP_TRAJ_NWARP, m.A);
		if (cta_bytes(eng->nwarp) <= 96 * 1024) { eng->sync_mode = SYNC_CTA; eng->smem_traj = cta_bytes(eng->nwarp); }
		else eng->sync_mode = SYNC_GRID;
	}
	if (eng->dense_m) eng->sync_mode = SYNC_GRID;           /* one launch per step, between the GEMMs */
	if (force && !eng->dense_m && m.ndcc == 0) {
		int f = atoi(force);
		if (f == SYNC_GRID) eng->sync_mode = SYNC_GRID;
		else if (f == SYNC_CTA && m.M_nnz == 0) { eng->sync_mode = SYNC_CTA; eng->nwarp = std::min(BEEP_TRAJ_NWARP, m.A); eng->smem_traj = cta_bytes(eng->nwarp); }
	}
	if (m.ndcc > 0 && m.A > 1) {
		/* a demographic change (state.cc:339-376) sums the population over ITS areas every step: the particle must sit
		 * in one CTA, whatever the mixing matrix says (and whatever a test forces) */
		eng->nwarp = std::min(BEEP_TRAJ_NWARP, m.A);
		if (cta_bytes(eng->nwarp) > 96 * 1024) return fail(BEEP_ERR_UNSUPPORTED, "democat_change over %d areas: the particle does not fit one CTA (%zu B of shared memory)", m.A, cta_bytes(eng->nwarp));
		eng->sync_mode = SYNC_CTA; eng->smem_traj = cta_bytes(eng->nwarp);
		eng->dense_m = false;
	}
	if (eng->sync_mode == SYNC_GRID) {
		eng->nwarp = BEEP_TRAJ_NWARP;
		eng->smem_traj = (common + eng->nwarp * (CDp + work)) * sizeof(double);
	}
	if ((long long)m.A * m.G >= (1ll << 28)) return fail(BEEP_ERR_UNSUPPORTED, "too many areas (A*G must be below 2^28)");
	if (eng->smem_traj > (size_t)eng->smem_optin) return fail(BEEP_ERR_UNSUPPORTED, "tile needs %zu B of shared memory (> %d)", eng->smem_traj, eng->smem_optin);
	for (int mode = 0; mode < 3; mode++) for (int keep = 0; keep < 2; keep++) {
		const void *fn = traj_fn(mode, eng->sync_mode, keep != 0);
		CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eng->smem_traj));
	}
	if (eng->sync_mode == SYNC_GRID) {
		int coop = 0;
		CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, eng->device));
		if (!coop) return fail(BEEP_ERR_UNSUPPORTED, "device lacks cooperative launch");
		int nb = 0, nbmin = 1 << 30;
		for (int mode = 0; mode < 3; mode++) for (int keep = 0; keep < 2; keep++) {
			CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, traj_fn(mode, SYNC_GRID, keep != 0), eng->nwarp * 32, eng->smem_traj));
			nbmin = std::min(nbmin, nb);
		}
		if (nbmin < 1) return fail(BEEP_ERR_UNSUPPORTED, "trajectory kernel does not fit an SM");
		eng->coop_grid = nbmin * eng->nsm;
		/* the chunk kernel for models of the common structure: NT tiles per warp and grab */
		eng->chunk_ok = false;
		const char *off = getenv("BEEP_NO_CHUNK");
		if (eng->chunk_model && !eng->dense_m && !(off && atoi(off))) {
			int nt = 4;                              /* measured: 4 tiles x 24 warps/SM beats 8 x 16 (profiles/README.md) */
			if (const char *e = getenv("BEEP_CHUNK_NT")) nt = (atoi(e) == 8) ? 8 : 4;
			if (m.A < 8) nt = 4;
			eng->chunk_nt = nt;
			eng->chunk_ls = chunk_sm_layout(nt, m.G, L.Gs, L.JD32, m.CDp4);
			eng->chunk_smem = common * sizeof(double) + (size_t)CHUNK_NWARP * eng->chunk_ls.per;
			if (eng->chunk_smem <= (size_t)eng->smem_optin) {
				int nbc = 1 << 30;
				for (int mode = 0; mode < 3; mode++) for (int keep = 0; keep < 2; keep++) {
					const void *fn = chunk_fn(mode, nt, keep != 0);
					CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eng->chunk_smem));
					CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, CHUNK_NWARP * 32, eng->chunk_smem));
					nbc = std::min(nbc, nb);
				}
				if (nbc >= 1) { eng->chunk_ok = true; eng->chunk_grid = nbc * eng->nsm; }
			}
		}
	}
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* batches                                                                   */
/* ------------------------------------------------------------------------ */
template <typename Tp>
static int balloc(BeepBatch *bt, Tp **out, size_t n)
{
	Tp *d = nullptr;
	size_t bytes = std::max<size_t>(n, 1) * sizeof(Tp);
	cudaError_t e = cudaMalloc((void **)&d, bytes);
	if (e != cudaSuccess) return fail(BEEP_ERR_CUDA, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
	bt->dev_allocs.push_back(d);
	bt->bytes += (long long)bytes;
	*out = d;
	return BEEP_OK;
}
#define BA(ptr, n) do { int rc_ = balloc(bt, &(ptr), (size_t)(n)); if (rc_) { beep_batch_destroy(bt); return rc_; } } while (0)

extern "C" void beep_batch_destroy(BeepBatch *bt)
{
	if (!bt) return;
	cudaSetDevice(bt->eng->device);
	if (bt->stream) cudaStreamSynchronize(bt->stream);
	for (void *p : bt->dev_allocs) cudaFree(p);
	if (bt->h_pinned) cudaFreeHost(bt->h_pinned);
	if (bt->ev0) cudaEventDestroy(bt->ev0);
	if (bt->ev1) cudaEventDestroy(bt->ev1);
	if (bt->stream) cudaStreamDestroy(bt->stream);
	delete bt;
}

extern "C" int beep_batch_create(BeepEngine *eng, int P, unsigned flags, BeepBatch **out)
{
	if (!eng || !out || P <= 0) return fail(BEEP_ERR_INVALID, "bad argument to beep_batch_create");
	if ((long long)P * eng->dm.A >= (1ll << 30) || P >= (1 << 22)) return fail(BEEP_ERR_UNSUPPORTED, "P*A must be below 2^30 tiles and P below 2^22");
	*out = nullptr;
	CU(cudaSetDevice(eng->device));
	BeepBatch *bt = new BeepBatch();
	bt->eng = eng;
	const DevModel &m = eng->dm;
	DevBatch &b = bt->db;
	memset(&b, 0, sizeof b);
	b.P = P; b.flags = flags;
	cudaError_t e = cudaStreamCreateWithFlags(&bt->stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) { delete bt; return fail(BEEP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
	cudaEventCreate(&bt->ev0); cudaEventCreate(&bt->ev1);
	const size_t TAJD = (size_t)m.T * m.A * m.J * m.D;
	BA(b.theta, (size_t)P * m.nparam); BA(b.pd, (size_t)P * m.nparam);
	BA(b.transrate, (size_t)P * m.J * m.D); BA(b.disc_spline, (size_t)P * m.nspline * m.T); BA(b.sus, (size_t)P * m.D);
	BA(b.areafactor, (size_t)P * m.period * m.A); BA(b.betaR, (size_t)P * m.S * m.nR * m.T); BA(b.infco, (size_t)P * m.C);
	/* stochastic batches of chunk-kernel models keep integer populations (int32, half the per-step traffic) */
	if (eng->chunk_ok && !(flags & BEEP_KEEP_POP)) BA(b.popi, (size_t)P * m.A * m.CDp4);
	else BA(b.pop_cur, (size_t)P * m.A * m.CDp);
	BA(b.scratchA, (size_t)P * m.A);
	if (eng->sync_mode == SYNC_GRID) BA(b.dinf, 2 * (size_t)P * m.S * m.A * m.G);
	if (eng->dense_m) BA(b.imap, (size_t)P * m.S * m.A * m.G);
	if ((flags & BEEP_KEEP_TRANSNUM) && m.stochastic) BA(b.tn, (size_t)P * TAJD);
	if ((flags & BEEP_KEEP_TRANSMEAN) || ((flags & BEEP_KEEP_TRANSNUM) && !m.stochastic)) { BA(b.tm, (size_t)P * TAJD); b.flags |= BEEP_KEEP_TRANSMEAN; }
	if (flags & BEEP_KEEP_POP) BA(b.poptraj, (size_t)P * m.T * m.A * m.C * m.D);
	BA(b.obs_partial, (size_t)P * std::max(m.nslot, 1)); BA(b.obs_val, (size_t)P * std::max(m.nobs, 1));
	BA(b.L, P); BA(b.EFtab, (size_t)P * std::max(m.ntable, 1)); BA(b.Li, P);
	BA(b.stream_id, P); BA(b.infmask, P);
	BA(b.crate, (size_t)P * (((size_t)m.J * m.D + 31) & ~(size_t)31) + 32);        /* + one warp pass: the chunk kernel requests a pass ahead without a bound test */
	BA(b.etab, (size_t)P * std::max(m.nefoi, 1) * m.T); BA(b.inflist, (size_t)P * (m.C + 1));
	if (eng->sync_mode == SYNC_GRID) BA(b.sched, ((size_t)P + 1) * 16);
	else if (eng->sync_mode == SYNC_NONE) BA(b.sched, 16);              /* the tile counter of the persistent warps */
	BA(bt->d_scratch, (size_t)P * std::max<size_t>(k5_chunks((size_t)m.T * m.A * m.J * m.D), (size_t)OBS_CHUNKS * (1 + std::max(m.ntable, 0))));
	BA(bt->d_map, 2 * (size_t)P); BA(bt->d_flag, P);
	{
		std::vector<unsigned long long> ids(P);
		for (int i = 0; i < P; i++) ids[i] = (unsigned long long)i;
		cudaMemcpy(b.stream_id, ids.data(), P * sizeof(unsigned long long), cudaMemcpyHostToDevice);
		cudaMemset(b.obs_partial, 0, (size_t)P * std::max(m.nslot, 1) * sizeof(double));
		cudaMemset(b.obs_val, 0, (size_t)P * std::max(m.nobs, 1) * sizeof(double));
	}
	bt->h_pinned_n = (size_t)P * std::max(std::max(m.nparam, m.ntable), 1);
	e = cudaMallocHost((void **)&bt->h_pinned, bt->h_pinned_n * sizeof(double));
	if (e != cudaSuccess) { beep_batch_destroy(bt); return fail(BEEP_ERR_CUDA, "cudaMallocHost: %s", cudaGetErrorString(e)); }
	*out = bt;
	return BEEP_OK;
}

extern "C" int beep_batch_size(const BeepBatch *bt) { return bt ? bt->db.P : 0; }
extern "C" int64_t beep_batch_bytes(const BeepBatch *bt) { return bt ? bt->bytes : 0; }
extern "C" void *beep_batch_stream(BeepBatch *bt) { return bt ? (void *)bt->stream : nullptr; }
extern "C" int64_t beep_launch_count(const BeepEngine *eng) { return eng ? eng->launches : 0; }
extern "C" int beep_engine_sync_scope(const BeepEngine *eng) { return eng ? eng->sync_mode : -1; }
extern "C" int beep_engine_dims(const BeepEngine *eng, int *T, int *A, int *C, int *J, int *D, int *nparam, int *nobs, int *ntable)
{
	if (!eng) return fail(BEEP_ERR_INVALID, "null engine");
	const DevModel &m = eng->dm;
	if (T) *T = m.T; if (A) *A = m.A; if (C) *C = m.C; if (J) *J = m.J; if (D) *D = m.D;
	if (nparam) *nparam = m.nparam; if (nobs) *nobs = m.nobs; if (ntable) *ntable = m.ntable;
	return BEEP_OK;
}
extern "C" int beep_engine_dense_mixing(const BeepEngine *eng) { return (eng && eng->dense_m) ? 1 : 0; }
extern "C" int beep_engine_chunk_tiles(const BeepEngine *eng) { return (eng && eng->chunk_ok) ? eng->chunk_nt : 0; }
extern "C" int beep_sync(BeepBatch *bt)
{
	if (!bt) return fail(BEEP_ERR_INVALID, "null batch");
	CU(cudaSetDevice(bt->eng->device));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}
extern "C" double beep_last_simulate_ms(const BeepBatch *bt)
{
	if (!bt || !bt->have_sim_time) return -1;
	cudaSetDevice(bt->eng->device);
	if (cudaEventSynchronize(bt->ev1) != cudaSuccess) return -1;
	float ms = 0;
	if (cudaEventElapsedTime(&ms, bt->ev0, bt->ev1) != cudaSuccess) return -1;
	return ms;
}

/* ------------------------------------------------------------------------ */
/* K0                                                                        */
/* ------------------------------------------------------------------------ */
static int launch_set_param(BeepBatch *bt)
{
	BeepEngine *eng = bt->eng;
	const DevModel &m = eng->dm;
	const int P = bt->db.P;
	k_set_param<<<P, K0_THREADS, 0, bt->stream>>>(m, bt->db);
	{
		/* beta: one warp per particle, as many particles per CTA as keep >= 4 CTAs' worth of shared memory per SM */
		const size_t per = k_beta_smem(m.Dq, m.G, m.nspline, m.nmm) * sizeof(double);
		int nw = (int)std::max<size_t>(1, std::min<size_t>(K0_BETA_WARPS, (48 * 1024) / per));
		const size_t smem = per * nw;
		if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_beta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_beta<<<(P + nw - 1) / nw, nw * 32, smem, bt->stream>>>(m, bt->db, nw);
		eng->launches++;
	}
	k_areafactor_fill<<<dim3(P, (m.A + 127) / 128), 128, 0, bt->stream>>>(m, bt->db);
	k_areafactor_sum<<<(P + 63) / 64, 64, 0, bt->stream>>>(m, bt->db);
	size_t naf = (size_t)m.period * m.A;
	k_areafactor_scale<<<dim3(P, (unsigned)std::min<size_t>((naf + 255) / 256, 1024)), 256, 0, bt->stream>>>(m, bt->db);
	eng->launches += 4;
	CU(cudaGetLastError());
	bt->theta_dirty = false;
	return BEEP_OK;
}
static void all_slots_have_theta(BeepBatch *bt) { bt->slot_theta.assign(bt->db.P, 1); bt->have_params = true; }

extern "C" int beep_set_params_dev(BeepBatch *bt, const double *theta_dev)
{
	if (!bt || !theta_dev) return fail(BEEP_ERR_INVALID, "null argument");
	CU(cudaSetDevice(bt->eng->device));
	CU(cudaMemcpyAsync(bt->db.theta, theta_dev, (size_t)bt->db.P * bt->eng->dm.nparam * sizeof(double), cudaMemcpyDeviceToDevice, bt->stream));
	all_slots_have_theta(bt);
	return launch_set_param(bt);
}

extern "C" int beep_set_params(BeepBatch *bt, const double *theta_host)
{
	if (!bt || !theta_host) return fail(BEEP_ERR_INVALID, "null argument");
	CU(cudaSetDevice(bt->eng->device));
	size_t n = (size_t)bt->db.P * bt->eng->dm.nparam;
	/* stage through the batch's pinned buffer so the copy is asynchronous and the caller's array is free on return */
	CU(cudaStreamSynchronize(bt->stream));
	memcpy(bt->h_pinned, theta_host, n * sizeof(double));
	CU(cudaMemcpyAsync(bt->db.theta, bt->h_pinned, n * sizeof(double), cudaMemcpyHostToDevice, bt->stream));
	all_slots_have_theta(bt);
	return launch_set_param(bt);
}

/* ------------------------------------------------------------------------ */
/* K1-K3                                                                     */
/* ------------------------------------------------------------------------ */
static int launch_traj(BeepBatch *bt, int mode, int ti, int tf, uint64_t seed, const BeepBatch *init, const unsigned char *d_mask)
{
	BeepEngine *eng = bt->eng;
	const DevModel &m = eng->dm;
	if (bt->theta_dirty) {
		/* State::initialise_from_particle starts with set_param (state.cc:296): K0 from the parameters on the device */
		int rc = launch_set_param(bt);
		if (rc) return rc;
	}
	TrajArgs a;
	a.m = m; a.b = bt->db; a.ti = ti; a.tf = tf; a.seed = seed;
	for (int r = 0; r < 10; r++) { a.rk[2 * r] = (unsigned)seed + (unsigned)r * 0x9E3779B9u; a.rk[2 * r + 1] = (unsigned)(seed >> 32) + (unsigned)r * 0xBB67AE85u; }
	a.init_tn = init ? init->db.tn : nullptr; a.init_tm = init ? init->db.tm : nullptr; a.simu_mask = d_mask;
	a.nwarp = eng->nwarp; a.chunk = 1; a.invG = 0; memset(&a.Ls, 0, sizeof a.Ls);
	a.invA = 1.0f / (float)m.A; a.L = traj_sm_layout(m.G, m.J, m.D, traj_nbr_doubles(m));
	/* tile_step_fast: one strain, dp == age, several ages, stochastic, no population trajectory */
	a.fast = (m.ndcc == 0 && !m.has_mod && m.S == 1 && m.age_fast && m.Dq == m.G && m.G > 1 && m.stochastic && !(bt->db.flags & BEEP_KEEP_POP)) ? 1 : 0;
	if (const char *nf = getenv("BEEP_TRAJ_NOFAST")) if (atoi(nf)) a.fast = 0;
	for (int tr = 0; tr < 32; tr++) { a.tfrom[tr] = tr < m.J ? eng->h_from[tr] : 0; a.tto[tr] = tr < m.J ? eng->h_to[tr] : 0; a.tkind[tr] = tr < m.J ? eng->h_kind[tr] : 2; }
	const int P = bt->db.P;
	const bool chunk = bt->db.popi != nullptr;
	if (ti == 0) {
		size_t n = (size_t)P * m.A * m.CDp;
		int nb = (int)std::min<size_t>((n + 255) / 256, (size_t)eng->nsm * 8);
		if (chunk) k_pop_reset_i32<<<nb, 256, 0, bt->stream>>>(m, bt->db);
		else k_pop_reset<<<nb, 256, 0, bt->stream>>>(m, bt->db);
		size_t ns = (size_t)P * std::max(m.nslot, 1);
		k_zero<<<(int)std::min<size_t>((ns + 255) / 256, (size_t)eng->nsm * 8), 256, 0, bt->stream>>>(bt->db.obs_partial, ns);
		eng->launches += 2;
	}
	CU(cudaEventRecord(bt->ev0, bt->stream));
	const bool keep = (bt->db.flags & (BEEP_KEEP_TRANSNUM | BEEP_KEEP_TRANSMEAN | BEEP_KEEP_POP)) != 0;
	a.imap_pre = nullptr;
	if (eng->dense_m) {
		/* dense mixing matrix: per step, Imap of ALL particles by one DMMA GEMM (the matrix streams once per step
		 * for the whole batch), then one step of the grid-spanning kernel that reads it */
		const void *fn1 = traj_fn(mode, SYNC_GRID, keep);
		void *args1[] = {(void *)&a};
		const long long ntile = (long long)P * m.A;
		const int grid = (int)std::min<long long>(eng->coop_grid, (ntile + eng->nwarp - 1) / eng->nwarp);
		const long long nw = (long long)grid * eng->nwarp;
		a.chunk = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(16, m.A / 4), std::max<long long>(mode == MODE_SIM ? 2 : 1, ntile / (nw * 4))));
		const size_t dsz = (size_t)P * m.S * m.A * m.G;
		const int NB = P * m.S, N = NB * m.G;
		k_dinf_init<<<(int)std::min<size_t>((dsz + 255) / 256, (size_t)eng->nsm * 8), 256, 0, bt->stream>>>(m, bt->db, bt->db.dinf + (ti & 1) * dsz);
		eng->launches++;
		a.imap_pre = bt->db.imap;
		/* the launch-bound step loop (three launches per step) is captured into a CUDA graph and replayed as one
		 * submission; if the capture is refused the launches go out one by one */
		auto enqueue = [&](cudaStream_t st) -> cudaError_t {
			for (int t = ti; t < tf; t++) {
				k_imap_dmma<<<dim3((N + IMD_TN - 1) / IMD_TN, (m.A + IMD_TM - 1) / IMD_TM), 128, 0, st>>>(m.Mt_dense, bt->db.dinf + (t & 1) * dsz, bt->db.imap, m.A, m.G, NB);
				k_sched_init<<<(P + 127) / 128, 128, 0, st>>>(bt->db.sched, P, t);
				a.ti = t; a.tf = t + 1;
				cudaError_t e = cudaLaunchCooperativeKernel(fn1, dim3(grid), dim3(eng->nwarp * 32), args1, eng->smem_traj, st);
				if (e != cudaSuccess) return e;
			}
			return cudaGetLastError();
		};
		bool graphed = false;
		if (tf - ti > 1 && !getenv("BEEP_NO_GRAPH")) {
			cudaGraph_t graph = nullptr;
			cudaGraphExec_t exec = nullptr;
			if (cudaStreamBeginCapture(bt->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
				cudaError_t e1 = enqueue(bt->stream);
				cudaError_t e2 = cudaStreamEndCapture(bt->stream, &graph);
				if (e1 == cudaSuccess && e2 == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
					graphed = cudaGraphLaunch(exec, bt->stream) == cudaSuccess;
					CU(cudaStreamSynchronize(bt->stream));         /* the graph's argument copies live until it has run */
				}
				if (exec) cudaGraphExecDestroy(exec);
				if (graph) cudaGraphDestroy(graph);
			}
			cudaGetLastError();
		}
		if (!graphed) CU(enqueue(bt->stream));
		eng->launches += 3ll * (tf - ti);
		CU(cudaEventRecord(bt->ev1, bt->stream));
		bt->have_sim_time = true;
		bt->cur_t = tf;
		return BEEP_OK;
	}
	const void *fn = traj_fn(mode, eng->sync_mode, keep);
	void *args[] = {(void *)&a};
	if (chunk) {
		a.Ls = eng->chunk_ls; a.invG = (65536u + m.G - 1) / m.G; a.chunk = eng->chunk_nt;
		long long nchunk = (long long)P * ((m.A + eng->chunk_nt - 1) / eng->chunk_nt);
		int grid = (int)std::min<long long>(eng->chunk_grid, (nchunk + CHUNK_NWARP - 1) / CHUNK_NWARP);
		k_sched_init<<<(P + 127) / 128, 128, 0, bt->stream>>>(bt->db.sched, P, ti);
		eng->launches++;
		CU(cudaLaunchCooperativeKernel(chunk_fn(mode, eng->chunk_nt, keep), dim3(grid), dim3(CHUNK_NWARP * 32), args, eng->chunk_smem, bt->stream));
	} else if (eng->sync_mode == SYNC_NONE) {
		/* persistent warps: as many CTAs as are resident at once (or fewer, if there are fewer tiles) */
		long long ntile = (long long)P * m.A;
		int grid = (int)std::min<long long>((ntile + eng->nwarp - 1) / eng->nwarp, (long long)eng->none_resident);
		CU(cudaMemsetAsync(bt->db.sched, 0, sizeof(unsigned long long), bt->stream));
		CU(cudaLaunchKernel(fn, dim3(grid), dim3(eng->nwarp * 32), args, eng->smem_traj, bt->stream));
	} else if (eng->sync_mode == SYNC_CTA) {
		CU(cudaLaunchKernel(fn, dim3(P), dim3(eng->nwarp * 32), args, eng->smem_traj, bt->stream));
	} else {
		long long ntile = (long long)P * m.A;
		int grid = (int)std::min<long long>(eng->coop_grid, (ntile + eng->nwarp - 1) / eng->nwarp);
		/* dynamic tile scheduling: about 4 grabs per warp and sweep over the particles (measured: 8 tiles per
		 * grab at 10 000 areas x 16 particles, 2 at x 4), at least 2 and at most 16 tiles per grab, and at
		 * least 4 grabs per particle-step */
		const long long nw = (long long)grid * eng->nwarp;
		a.chunk = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(16, m.A / 4), std::max<long long>(mode == MODE_SIM ? 2 : 1, ntile / (nw * 4))));     /* MBP / replay tiles are heavier: finer grabs when tiles are scarce */
		if (const char *ch = getenv("BEEP_TRAJ_CHUNK")) a.chunk = std::max(1, atoi(ch));
		k_sched_init<<<(P + 127) / 128, 128, 0, bt->stream>>>(bt->db.sched, P, ti);
		eng->launches++;
		CU(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(eng->nwarp * 32), args, eng->smem_traj, bt->stream));
	}
	eng->launches++;
	CU(cudaEventRecord(bt->ev1, bt->stream));
	bt->have_sim_time = true;
	bt->cur_t = tf;
	return BEEP_OK;
}

static int set_stream_ids(BeepBatch *bt, const uint64_t *ids)
{
	if (!ids) return BEEP_OK;
	CU(cudaStreamSynchronize(bt->stream));
	CU(cudaMemcpyAsync(bt->db.stream_id, ids, (size_t)bt->db.P * sizeof(uint64_t), cudaMemcpyHostToDevice, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

extern "C" int beep_simulate(BeepBatch *bt, int ti, int tf, uint64_t seed, const uint64_t *stream_id_host)
{
	if (!bt) return fail(BEEP_ERR_INVALID, "null batch");
	const DevModel &m = bt->eng->dm;
	if (!bt->have_params) return fail(BEEP_ERR_STATE, "beep_simulate before beep_set_params");
	if (ti < 0 || tf > m.T || ti >= tf) return fail(BEEP_ERR_INVALID, "bad step range [%d,%d)", ti, tf);
	if (ti != 0 && ti != bt->cur_t) return fail(BEEP_ERR_STATE, "simulate(%d,..) but the batch stands at step %d", ti, bt->cur_t);
	CU(cudaSetDevice(bt->eng->device));
	int rc = set_stream_ids(bt, stream_id_host);
	if (rc) return rc;
	return launch_traj(bt, MODE_SIM, ti, tf, seed, nullptr, nullptr);
}

extern "C" int beep_mbp(BeepBatch *init, BeepBatch *prop, const uint8_t *simu_mask_host, int inf_update, uint64_t seed, const uint64_t *stream_id_host)
{
	(void)inf_update;
	if (!init || !prop || init == prop) return fail(BEEP_ERR_INVALID, "beep_mbp needs two distinct batches");
	if (init->eng != prop->eng || init->db.P != prop->db.P) return fail(BEEP_ERR_INVALID, "init and prop batches differ in engine or size");
	const DevModel &m = prop->eng->dm;
	if (!m.stochastic) return fail(BEEP_ERR_UNSUPPORTED, "model-based proposals need stochastic dynamics");
	if (!init->db.tn || !init->db.tm) return fail(BEEP_ERR_STATE, "init batch must keep transnum and transmean");
	if (!prop->have_params) return fail(BEEP_ERR_STATE, "beep_mbp before beep_set_params on the proposal batch");
	CU(cudaSetDevice(prop->eng->device));
	int rc = set_stream_ids(prop, stream_id_host);
	if (rc) return rc;
	unsigned char *d_mask = nullptr;
	if (simu_mask_host) {
		size_t n = (size_t)prop->db.P * m.T * m.A;
		if (prop->d_mask_n < n) {
			rc = balloc(prop, &prop->d_mask, n);
			if (rc) return rc;
			prop->d_mask_n = n;
		}
		CU(cudaStreamSynchronize(prop->stream));
		CU(cudaMemcpyAsync(prop->d_mask, simu_mask_host, n, cudaMemcpyHostToDevice, prop->stream));
		d_mask = prop->d_mask;
	}
	/* order after everything queued on the init batch */
	CU(cudaStreamSynchronize(init->stream));
	return launch_traj(prop, MODE_MBP, 0, m.T, seed, init, d_mask);
}

extern "C" int beep_replay(BeepBatch *bt)
{
	if (!bt) return fail(BEEP_ERR_INVALID, "null batch");
	const DevModel &m = bt->eng->dm;
	if (!bt->have_params) return fail(BEEP_ERR_STATE, "beep_replay before parameters were set");
	if (m.stochastic ? !bt->db.tn : !bt->db.tm) return fail(BEEP_ERR_STATE, "beep_replay needs a batch that keeps transnum");
	CU(cudaSetDevice(bt->eng->device));
	return launch_traj(bt, MODE_REPLAY, 0, m.T, 0, nullptr, nullptr);
}

/* The same for the steps [ti, tf) only: afterwards the batch stands at step tf and beep_simulate(batch, tf, ...) carries on from
 * the replayed population -- State::load followed by State::simulate(pred_start, ndivision) of mode=prediction
 * (Simulate::model_modification, simulate.cc:141-165) */
extern "C" int beep_replay_range(BeepBatch *bt, int ti, int tf)
{
	if (!bt) return fail(BEEP_ERR_INVALID, "null batch");
	const DevModel &m = bt->eng->dm;
	if (!bt->have_params) return fail(BEEP_ERR_STATE, "beep_replay_range before parameters were set");
	if (m.stochastic ? !bt->db.tn : !bt->db.tm) return fail(BEEP_ERR_STATE, "beep_replay_range needs a batch that keeps transnum");
	if (ti < 0 || tf > m.T || ti >= tf) return fail(BEEP_ERR_INVALID, "bad step range [%d,%d)", ti, tf);
	if (ti != 0 && ti != bt->cur_t) return fail(BEEP_ERR_STATE, "replay(%d,..) but the batch stands at step %d", ti, bt->cur_t);
	CU(cudaSetDevice(bt->eng->device));
	if (bt->theta_dirty) { int rc = launch_set_param(bt); if (rc) return rc; }
	return launch_traj(bt, MODE_REPLAY, ti, tf, 0, nullptr, nullptr);
}

/* ------------------------------------------------------------------------ */
/* K4 / K5                                                                   */
/* ------------------------------------------------------------------------ */
extern "C" int beep_obs_loglik_dev(BeepBatch *bt, int section, double *L_dev, double *EF_dev)
{
	if (!bt) return fail(BEEP_ERR_INVALID, "null batch");
	BeepEngine *eng = bt->eng;
	if (section >= eng->dm.nsection) return fail(BEEP_ERR_INVALID, "section %d out of range", section);
	CU(cudaSetDevice(eng->device));
	const int P = bt->db.P, wt = EF_dev ? 1 : 0, nq = wt ? 1 + eng->dm.ntable : 1;
	/* chunks of at least a block's worth of observations; small observation sets get small blocks */
	const int nobs = eng->dm.nobs;
	const int nch = std::max(1, std::min(OBS_CHUNKS, (nobs + RED_THREADS - 1) / RED_THREADS));
	const int nth = std::min(RED_THREADS, std::max(32, ((nobs + nch - 1) / nch + 31) & ~31));
	k_obs_final<<<dim3(P, nch), nth, 0, bt->stream>>>(eng->dm, bt->db, section, wt, bt->d_scratch);
	k_obs_reduce<<<(P * nq + 127) / 128, 128, 0, bt->stream>>>(eng->dm, P, wt, nch, bt->d_scratch, L_dev ? L_dev : bt->db.L, EF_dev);
	eng->launches += 2;
	CU(cudaGetLastError());
	return BEEP_OK;
}

extern "C" int beep_obs_loglik(BeepBatch *bt, int section, double *L_host, double *EF_table_host)
{
	if (!bt || !L_host) return fail(BEEP_ERR_INVALID, "null argument");
	int rc = beep_obs_loglik_dev(bt, section, bt->db.L, EF_table_host ? bt->db.EFtab : nullptr);
	if (rc) return rc;
	const int P = bt->db.P, nt = bt->eng->dm.ntable;
	CU(cudaMemcpyAsync(L_host, bt->db.L, P * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
	if (EF_table_host) CU(cudaMemcpyAsync(EF_table_host, bt->db.EFtab, (size_t)P * nt * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	for (int p = 0; p < P; p++) if (std::isnan(L_host[p])) return fail(BEEP_ERR_NUMERIC, "observation log-likelihood of particle %d is NaN (obsmodel.cc:30)", p);
	return BEEP_OK;
}

extern "C" int beep_obs_values(BeepBatch *bt, double *out)
{
	if (!bt || !out) return fail(BEEP_ERR_INVALID, "null argument");
	int rc = beep_obs_loglik_dev(bt, -1, bt->db.L, nullptr);
	if (rc) return rc;
	CU(cudaMemcpyAsync(out, bt->db.obs_val, (size_t)bt->db.P * bt->eng->dm.nobs * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

extern "C" int beep_num_sections(const BeepEngine *eng) { return eng ? eng->dm.nsection : 0; }
extern "C" int beep_section_range(const BeepEngine *eng, int s, int *ti, int *tf)
{
	if (!eng || s < 0 || s >= eng->dm.nsection) return fail(BEEP_ERR_INVALID, "section out of range");
	if (ti) *ti = eng->section_ti[s];
	if (tf) *tf = eng->section_tf[s];
	return BEEP_OK;
}

extern "C" int beep_latent_loglik(BeepBatch *bt, double *Li_host)
{
	if (!bt || !Li_host) return fail(BEEP_ERR_INVALID, "null argument");
	BeepEngine *eng = bt->eng;
	const DevModel &m = eng->dm;
	if (!bt->db.tm || (m.stochastic && !bt->db.tn)) return fail(BEEP_ERR_STATE, "beep_latent_loglik needs a batch that keeps transnum and transmean");
	CU(cudaSetDevice(eng->device));
	const int P = bt->db.P;
	const int nch = k5_chunks((size_t)m.T * m.A * m.J * m.D);
	k_latent_partial<<<dim3(P, nch), RED_THREADS, 0, bt->stream>>>(m, bt->db, bt->d_scratch);
	k_latent_final<<<(P + 127) / 128, 128, 0, bt->stream>>>(P, nch, bt->d_scratch, bt->db.Li);
	eng->launches += 2;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(Li_host, bt->db.Li, P * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

/* State::likelihood_approx (state_mvn_approx.cc:11-91) for every particle of the batch, at inverse temperature invT */
extern "C" int beep_likelihood_approx(BeepBatch *bt, double invT, double *L_host)
{
	if (!bt || !L_host) return fail(BEEP_ERR_INVALID, "null argument");
	BeepEngine *eng = bt->eng;
	const DevModel &m = eng->dm;
	if (!eng->approx_ok) return fail(BEEP_ERR_UNSUPPORTED, "beep_likelihood_approx: %s", eng->approx_why.c_str());
	if (!bt->have_params) return fail(BEEP_ERR_STATE, "beep_likelihood_approx before beep_set_params");
	if (!(invT > 0)) return fail(BEEP_ERR_INVALID, "invT must be positive");
	CU(cudaSetDevice(eng->device));
	if (bt->theta_dirty) { int rc = launch_set_param(bt); if (rc) return rc; }
	const int P = bt->db.P;
	DevApprox x = eng->ax;
	const size_t n = x.n, nt = x.nt, nw = x.nwo, vs = std::max(n, nt);
	/* entries of the infection rows of the Jacobian: D ((nnz + A) ninf D + A), ninf <= C - 1 */
	const size_t jcap = (size_t)m.D * (((size_t)m.M_nnz + m.A) * (size_t)(m.C - 1) * m.D + m.A);
	/* the step vectors, and the covariance with the product R when they leave room for several CTAs per SM, live in shared memory */
	const size_t sm_vec_b = AP_NVEC * vs * sizeof(double), sm_mat_b = (n * n + nt * n) * sizeof(double);
	x.sm_vec = sm_vec_b <= 40 * 1024;
	x.sm_mat = x.sm_vec && sm_vec_b + sm_mat_b <= 72 * 1024;
	const size_t smem = (x.sm_vec ? sm_vec_b : 0) + (x.sm_mat ? sm_mat_b : 0);
	const size_t per = (size_t)(m.T + 1) * n + (x.sm_mat ? 0 : n * n + nt * n) + 6 * nw * nw + (x.sm_vec ? 0 : AP_NVEC * vs) + 1 + jcap + (jcap + 1) / 2;
	const size_t need = per * P * sizeof(double);
	if (bt->d_approx_n < need) {
		CU(cudaStreamSynchronize(bt->stream));
		if (bt->d_approx) {
			cudaFree(bt->d_approx);
			bt->dev_allocs.erase(std::remove(bt->dev_allocs.begin(), bt->dev_allocs.end(), (void *)bt->d_approx), bt->dev_allocs.end());
			bt->bytes -= (long long)bt->d_approx_n;
			bt->d_approx = nullptr; bt->d_approx_n = 0;
		}
		double *t = nullptr;
		cudaError_t e = cudaMalloc((void **)&t, need);
		if (e != cudaSuccess) return fail(BEEP_ERR_CUDA, "cudaMalloc(%zu bytes) for the approximate likelihood: %s", need, cudaGetErrorString(e));
		bt->dev_allocs.push_back(t);
		bt->bytes += (long long)need;
		bt->d_approx = t; bt->d_approx_n = need;
	}
	double *q = bt->d_approx;
	x.pop = q; q += (size_t)P * (m.T + 1) * n;
	if (!x.sm_mat) { x.covar = q; q += (size_t)P * n * n; x.R = q; q += (size_t)P * nt * n; }
	x.lu = q; q += (size_t)P * 6 * nw * nw;
	if (!x.sm_vec) { x.vec = q; q += (size_t)P * AP_NVEC * vs; }
	x.L = q; q += P;
	x.jval = q; q += (size_t)P * jcap;
	x.jcol = (int *)q;
	x.jcap = (long long)jcap;
	/* threads per particle: about four outputs of the n x n passes per thread, one to eight warps */
	int nth = (int)std::min<size_t>(AP_THREADS, std::max<size_t>(32, ((n * n / 4 + 31) / 32) * 32));
	x.logW = 0;
	while (x.logW < 5 && (1u << x.logW) < n) x.logW++;
	if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_likelihood_approx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_likelihood_approx<<<P, nth, smem, bt->stream>>>(m, bt->db, x, invT);
	eng->launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(L_host, x.L, P * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

/* 1 when the model qualifies for beep_likelihood_approx */
extern "C" int beep_engine_has_approx(const BeepEngine *eng) { return eng && eng->approx_ok ? 1 : 0; }

/* ObservationModel::get_graph_state (obsmodel.cc:113-170) of every particle: out_host [P][npoint], the graphs' points concatenated */
extern "C" int beep_graph_layout(const BeepEngine *eng, int *ngraph, int *npoint, int *point_off)
{
	if (!eng) return fail(BEEP_ERR_INVALID, "null engine");
	if (ngraph) *ngraph = eng->graphs_ok ? eng->gx.ngraph : 0;
	if (npoint) *npoint = eng->graphs_ok ? eng->gx.npoint : 0;
	if (point_off && eng->graphs_ok) for (size_t k = 0; k < eng->graph_point_off.size(); k++) point_off[k] = eng->graph_point_off[k];
	return BEEP_OK;
}

extern "C" int beep_graph_state(BeepBatch *bt, double *out_host)
{
	if (!bt || !out_host) return fail(BEEP_ERR_INVALID, "null argument");
	BeepEngine *eng = bt->eng;
	const DevModel &m = eng->dm;
	if (!eng->graphs_ok) return fail(BEEP_ERR_UNSUPPORTED, "the description carries no graphs");
	if (!bt->have_params) return fail(BEEP_ERR_STATE, "beep_graph_state before beep_set_params");
	if (!bt->db.tn && !bt->db.tm) return fail(BEEP_ERR_STATE, "beep_graph_state needs a batch that keeps transnum");
	if (eng->graphs_pop && !bt->db.poptraj) return fail(BEEP_ERR_STATE, "a graph plots populations: the batch must keep pop (BEEP_KEEP_POP)");
	CU(cudaSetDevice(eng->device));
	if (eng->graphs_obs) { int rc = beep_obs_loglik_dev(bt, -1, bt->db.L, nullptr); if (rc) return rc; }
	const int P = bt->db.P, np = eng->gx.npoint;
	const size_t need = (size_t)P * np * sizeof(double);
	if (bt->d_approx_n < need) {
		CU(cudaStreamSynchronize(bt->stream));
		if (bt->d_approx) {
			cudaFree(bt->d_approx);
			bt->dev_allocs.erase(std::remove(bt->dev_allocs.begin(), bt->dev_allocs.end(), (void *)bt->d_approx), bt->dev_allocs.end());
			bt->bytes -= (long long)bt->d_approx_n;
			bt->d_approx = nullptr; bt->d_approx_n = 0;
		}
		double *t = nullptr;
		cudaError_t e = cudaMalloc((void **)&t, need);
		if (e != cudaSuccess) return fail(BEEP_ERR_CUDA, "cudaMalloc(%zu bytes) for the graph states: %s", need, cudaGetErrorString(e));
		bt->dev_allocs.push_back(t);
		bt->bytes += (long long)need;
		bt->d_approx = t; bt->d_approx_n = need;
	}
	k_graph_state<<<dim3((unsigned)std::min((np + 7) / 8, 4096), (unsigned)P), 256, 0, bt->stream>>>(m, bt->db, eng->gx, bt->d_approx);
	eng->launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out_host, bt->d_approx, need, cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

/* Output::get_statistic (output.cc:3089-3123) per point over the P samples: mean and the 95 % credible interval by linear
 * interpolation of the sorted values -- what the reference writes to Posterior/state/<graph>.csv (output.cc:1109-1125) */
extern "C" int beep_graph_statistics(const double *gs_host, int P, int npoint, double *mean, double *CImin, double *CImax)
{
	if (!gs_host || P <= 0 || npoint <= 0 || !mean || !CImin || !CImax) return fail(BEEP_ERR_INVALID, "bad argument to beep_graph_statistics");
	std::vector<double> v(P);
	for (int i = 0; i < npoint; i++) {
		double sum = 0.0;
		for (int p = 0; p < P; p++) { v[p] = gs_host[(size_t)p * npoint + i]; sum += v[p]; }
		mean[i] = sum / P;
		std::sort(v.begin(), v.end());
		if (P >= 2) {
			unsigned k = (unsigned)((P - 1) * 0.025); double f = (P - 1) * 0.025 - k;
			CImin[i] = v[k] * (1 - f) + v[k + 1] * f;
			k = (unsigned)((P - 1) * 0.975); f = (P - 1) * 0.975 - k;
			CImax[i] = v[k] * (1 - f) + v[k + 1] * f;
		} else { CImin[i] = v[0]; CImax[i] = v[0]; }
	}
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* K6                                                                        */
/* ------------------------------------------------------------------------ */
static int permute_scratch(BeepBatch *bt, size_t need)
{
	if (bt->d_tmp_n >= need) return BEEP_OK;
	/* one scratch per batch, as large as the largest set of slabs ever moved: the smaller one is released first
	 * (stream-ordered after the copies that used it) and the batch's byte count follows */
	CU(cudaStreamSynchronize(bt->stream));
	if (bt->d_tmp) {
		cudaFree(bt->d_tmp);
		bt->dev_allocs.erase(std::remove(bt->dev_allocs.begin(), bt->dev_allocs.end(), (void *)bt->d_tmp), bt->dev_allocs.end());
		bt->bytes -= (long long)bt->d_tmp_n;
		bt->d_tmp = nullptr; bt->d_tmp_n = 0;
	}
	double *t = nullptr;
	cudaError_t e = cudaMalloc((void **)&t, need);
	if (e != cudaSuccess) return fail(BEEP_ERR_CUDA, "cudaMalloc(%zu bytes) for permute scratch: %s", need, cudaGetErrorString(e));
	bt->dev_allocs.push_back(t);
	bt->bytes += (long long)need;
	bt->d_tmp = t; bt->d_tmp_n = need;
	return BEEP_OK;
}

/* the per-particle tables of a list in two launches (gather the moved slots' sources, write them into place) */
struct SlabAdder {
	SlabList L;
	SlabAdder() { memset(&L, 0, sizeof L); }
	template <typename Tp> void add(Tp *p, size_t n) { if (p && n && L.n < 24) { L.ptr[L.n] = (void *)p; L.words[L.n] = (unsigned)(n * sizeof(Tp) / 4); L.off[L.n + 1] = L.off[L.n] + L.words[L.n]; L.n++; } }
};
static int gather_multi(BeepBatch *bt, const SlabList &L, const unsigned *d_map, int nm)
{
	if (L.n == 0 || nm == 0) return BEEP_OK;
	int rc = permute_scratch(bt, (size_t)L.off[L.n] * nm * 4);
	if (rc) return rc;
	unsigned wmax = 0;
	for (int s = 0; s < L.n; s++) wmax = std::max(wmax, L.words[s]);
	const unsigned nz = std::max(1u, std::min(64u, (wmax + 4095u) / 4096u));
	k_gather_multi<<<dim3(nm, L.n, nz), 256, 0, bt->stream>>>(L, (unsigned *)bt->d_tmp, d_map, nm);
	k_scatter_multi<<<dim3(nm, L.n, nz), 256, 0, bt->stream>>>(L, (const unsigned *)bt->d_tmp, d_map + nm, nm);
	bt->eng->launches += 2;
	CU(cudaGetLastError());
	return BEEP_OK;
}

/* arr[dst[k]] <- arr[src[k]] for the nm slots that change (d_map: nm sources, then nm destinations).  The sources are read
 * into a compact scratch first, so chains of copies are well defined (mpi.cc:81-165); slots that keep their particle cost
 * nothing, and the scratch holds nm slabs, not the whole array */
template <typename Tp>
static int gather_inplace(BeepBatch *bt, Tp *arr, size_t n, const unsigned *d_map, int nm)
{
	if (!arr || n == 0 || nm == 0) return BEEP_OK;
	int rc0 = permute_scratch(bt, (size_t)nm * n * sizeof(Tp));
	if (rc0) return rc0;
	Tp *tmp = (Tp *)bt->d_tmp;
	int gx = (int)std::min<size_t>((n + 255) / 256, 1024);
	k_gather_slab<Tp><<<dim3(nm, gx), 256, 0, bt->stream>>>(tmp, arr, d_map, n, nm);
	k_scatter_slab<Tp><<<dim3(nm, gx), 256, 0, bt->stream>>>(arr, tmp, d_map + nm, n);
	bt->eng->launches += 2;
	CU(cudaGetLastError());
	return BEEP_OK;
}

extern "C" int beep_permute(BeepBatch *bt, const uint32_t *src_of_host, int what)
{
	if (!bt || !src_of_host) return fail(BEEP_ERR_INVALID, "null argument");
	const DevModel &m = bt->eng->dm;
	const int P = bt->db.P;
	for (int p = 0; p < P; p++) if (src_of_host[p] != 0xFFFFFFFFu && src_of_host[p] >= (unsigned)P) return fail(BEEP_ERR_INVALID, "src_of[%d] = %u out of range", p, src_of_host[p]);
	CU(cudaSetDevice(bt->eng->device));
	int rc;
	if (what != BEEP_PERMUTE_SECTION_STATE && bt->theta_dirty && (rc = launch_set_param(bt))) return rc;   /* derived tables move with the particles */
	/* the slots that receive another particle: sources first, destinations after them */
	std::vector<unsigned> mv;
	for (int p = 0; p < P; p++) if (src_of_host[p] != 0xFFFFFFFFu && src_of_host[p] != (unsigned)p) mv.push_back(src_of_host[p]);
	const int nm = (int)mv.size();
	if (nm == 0) return BEEP_OK;
	for (int p = 0; p < P; p++) if (src_of_host[p] != 0xFFFFFFFFu && src_of_host[p] != (unsigned)p) mv.push_back((unsigned)p);
	CU(cudaStreamSynchronize(bt->stream));
	CU(cudaMemcpyAsync(bt->d_map, mv.data(), mv.size() * sizeof(unsigned), cudaMemcpyHostToDevice, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));                     /* mv is a local */
	const DevBatch &b = bt->db;
	if ((rc = gather_inplace(bt, b.pop_cur, (size_t)m.A * m.CDp, bt->d_map, nm))) return rc;
	if ((rc = gather_inplace(bt, b.popi, (size_t)m.A * m.CDp4, bt->d_map, nm))) return rc;
	if (what == BEEP_PERMUTE_SECTION_STATE) return BEEP_OK;      /* mpi.cc:286-288: pop (Imap/Idiag are re-derived from it) */
	const size_t TAJD = (size_t)m.T * m.A * m.J * m.D;
	{
		/* the parameters and the derived tables of K0, the observation sums: one pair of launches */
		SlabAdder sl;
		sl.add(b.theta, m.nparam); sl.add(b.pd, m.nparam); sl.add(b.transrate, (size_t)m.J * m.D); sl.add(b.disc_spline, (size_t)m.nspline * m.T);
		sl.add(b.sus, m.D); sl.add(b.areafactor, (size_t)m.period * m.A); sl.add(b.betaR, (size_t)m.S * m.nR * m.T); sl.add(b.infco, m.C);
		sl.add(b.infmask, 1); sl.add(b.etab, (size_t)m.nefoi * m.T); sl.add(b.crate, ((size_t)m.J * m.D + 31) & ~(size_t)31); sl.add(b.inflist, (size_t)m.C + 1);
		sl.add(b.obs_partial, m.nslot); sl.add(b.obs_val, m.nobs);
		if ((rc = gather_multi(bt, sl.L, bt->d_map, nm))) return rc;
	}
	if ((rc = gather_inplace(bt, b.tn, TAJD, bt->d_map, nm))) return rc;
	if ((rc = gather_inplace(bt, b.tm, TAJD, bt->d_map, nm))) return rc;
	if ((rc = gather_inplace(bt, b.poptraj, (size_t)m.T * m.A * m.C * m.D, bt->d_map, nm))) return rc;
	return BEEP_OK;
}

template <typename Tp>
static int copy_flagged(BeepBatch *dst, Tp *d, const Tp *s, size_t n)
{
	if (!d || !s || n == 0) return BEEP_OK;
	int gx = (int)std::min<size_t>((n + 255) / 256, 1024);
	k_copy_flagged<Tp><<<dim3(dst->db.P, gx), 256, 0, dst->stream>>>(d, s, dst->d_flag, n, dst->db.P);
	dst->eng->launches++;
	CU(cudaGetLastError());
	return BEEP_OK;
}

extern "C" int beep_accept(BeepBatch *init, BeepBatch *prop, const uint8_t *accept_host)
{
	if (!init || !prop || !accept_host || init == prop) return fail(BEEP_ERR_INVALID, "bad argument to beep_accept");
	if (init->eng != prop->eng || init->db.P != prop->db.P) return fail(BEEP_ERR_INVALID, "init and prop batches differ in engine or size");
	const DevModel &m = init->eng->dm;
	CU(cudaSetDevice(init->eng->device));
	if (prop->theta_dirty) { int rc0 = launch_set_param(prop); if (rc0) return rc0; }
	if (init->theta_dirty && init->have_params) { int rc0 = launch_set_param(init); if (rc0) return rc0; }
	CU(cudaStreamSynchronize(prop->stream));
	CU(cudaStreamSynchronize(init->stream));
	CU(cudaMemcpyAsync(init->d_flag, accept_host, init->db.P, cudaMemcpyHostToDevice, init->stream));
	const DevBatch &a = init->db, &b = prop->db;
	const size_t TAJD = (size_t)m.T * m.A * m.J * m.D;
	int rc;
	if ((rc = copy_flagged(init, a.theta, b.theta, m.nparam))) return rc;
	if ((rc = copy_flagged(init, a.pd, b.pd, m.nparam))) return rc;
	if ((rc = copy_flagged(init, a.transrate, b.transrate, (size_t)m.J * m.D))) return rc;
	if ((rc = copy_flagged(init, a.disc_spline, b.disc_spline, (size_t)m.nspline * m.T))) return rc;
	if ((rc = copy_flagged(init, a.sus, b.sus, m.D))) return rc;
	if ((rc = copy_flagged(init, a.areafactor, b.areafactor, (size_t)m.period * m.A))) return rc;
	if ((rc = copy_flagged(init, a.betaR, b.betaR, (size_t)m.S * m.nR * m.T))) return rc;
	if ((rc = copy_flagged(init, a.infco, b.infco, m.C))) return rc;
	if ((rc = copy_flagged(init, a.infmask, b.infmask, 1))) return rc;
	if ((rc = copy_flagged(init, a.etab, b.etab, (size_t)m.nefoi * m.T))) return rc;
	if ((rc = copy_flagged(init, a.crate, b.crate, ((size_t)m.J * m.D + 31) & ~(size_t)31))) return rc;
	if ((rc = copy_flagged(init, a.inflist, b.inflist, (size_t)m.C + 1))) return rc;
	if ((a.popi != nullptr) != (b.popi != nullptr)) return fail(BEEP_ERR_INVALID, "init and prop batches were created with different flags (population storage differs)");
	if ((rc = copy_flagged(init, a.pop_cur, b.pop_cur, (size_t)m.A * m.CDp))) return rc;
	if ((rc = copy_flagged(init, a.popi, b.popi, (size_t)m.A * m.CDp4))) return rc;
	if ((rc = copy_flagged(init, a.tn, b.tn, TAJD))) return rc;
	if ((rc = copy_flagged(init, a.tm, b.tm, TAJD))) return rc;
	if ((rc = copy_flagged(init, a.poptraj, b.poptraj, (size_t)m.T * m.A * m.C * m.D))) return rc;
	if ((rc = copy_flagged(init, a.obs_partial, b.obs_partial, m.nslot))) return rc;
	if ((rc = copy_flagged(init, a.obs_val, b.obs_val, m.nobs))) return rc;
	CU(cudaStreamSynchronize(init->stream));
	init->have_params = true;
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* particle import / export, read-back                                       */
/* ------------------------------------------------------------------------ */
static int slot_ok(const BeepBatch *bt, int slot)
{
	if (!bt) return fail(BEEP_ERR_INVALID, "null batch");
	if (slot < 0 || slot >= bt->db.P) return fail(BEEP_ERR_INVALID, "slot %d out of range", slot);
	return BEEP_OK;
}

extern "C" int beep_get_trajectory(BeepBatch *bt, int slot, int which, double *out, int64_t n)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	if (!out) return fail(BEEP_ERR_INVALID, "null output");
	const DevModel &m = bt->eng->dm;
	CU(cudaSetDevice(bt->eng->device));
	const size_t TAJD = (size_t)m.T * m.A * m.J * m.D, TACD = (size_t)m.T * m.A * m.C * m.D;
	if (which == BEEP_TRAJ_TRANSNUM && m.stochastic) {
		if (!bt->db.tn) return fail(BEEP_ERR_STATE, "batch does not keep transnum");
		if ((size_t)n != TAJD) return fail(BEEP_ERR_INVALID, "transnum has %zu elements, not %lld", TAJD, (long long)n);
		std::vector<int> tmp(TAJD);
		CU(cudaMemcpyAsync(tmp.data(), bt->db.tn + (size_t)slot * TAJD, TAJD * sizeof(int), cudaMemcpyDeviceToHost, bt->stream));
		CU(cudaStreamSynchronize(bt->stream));
		for (size_t i = 0; i < TAJD; i++) out[i] = (double)tmp[i];
		return BEEP_OK;
	}
	if (which == BEEP_TRAJ_TRANSMEAN || which == BEEP_TRAJ_TRANSNUM) {
		/* deterministic dynamics: transnum IS transmean (state.cc:627) */
		if (!bt->db.tm) return fail(BEEP_ERR_STATE, "batch does not keep transmean");
		if ((size_t)n != TAJD) return fail(BEEP_ERR_INVALID, "transmean has %zu elements, not %lld", TAJD, (long long)n);
		CU(cudaMemcpyAsync(out, bt->db.tm + (size_t)slot * TAJD, TAJD * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
		CU(cudaStreamSynchronize(bt->stream));
		return BEEP_OK;
	}
	if (which == BEEP_TRAJ_POP) {
		if (!bt->db.poptraj) return fail(BEEP_ERR_STATE, "batch does not keep pop");
		if ((size_t)n != TACD) return fail(BEEP_ERR_INVALID, "pop has %zu elements, not %lld", TACD, (long long)n);
		CU(cudaMemcpyAsync(out, bt->db.poptraj + (size_t)slot * TACD, TACD * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
		CU(cudaStreamSynchronize(bt->stream));
		return BEEP_OK;
	}
	return fail(BEEP_ERR_INVALID, "unknown trajectory selector %d", which);
}

/* integer populations (batches of the chunk kernel) cross the boundary as doubles like the others */
static int get_pop_i32(BeepBatch *bt, const int *src, size_t ntile, double *out)
{
	const DevModel &m = bt->eng->dm;
	const size_t CD = (size_t)m.C * m.D;
	std::vector<int> tmp(ntile * CD);
	CU(cudaMemcpy2DAsync(tmp.data(), CD * sizeof(int), src, (size_t)m.CDp4 * sizeof(int), CD * sizeof(int), ntile, cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	for (size_t i = 0; i < tmp.size(); i++) out[i] = (double)tmp[i];
	return BEEP_OK;
}

extern "C" int beep_get_pop(BeepBatch *bt, int slot, double *out, int64_t n)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	const DevModel &m = bt->eng->dm;
	size_t ACD = (size_t)m.A * m.C * m.D;
	if (!out || (size_t)n != ACD) return fail(BEEP_ERR_INVALID, "pop has %zu elements", ACD);
	CU(cudaSetDevice(bt->eng->device));
	const size_t CD = (size_t)m.C * m.D;
	if (bt->db.popi) return get_pop_i32(bt, bt->db.popi + (size_t)slot * m.A * m.CDp4, (size_t)m.A, out);
	CU(cudaMemcpy2DAsync(out, CD * sizeof(double), bt->db.pop_cur + (size_t)slot * m.A * m.CDp, (size_t)m.CDp * sizeof(double),
	                     CD * sizeof(double), m.A, cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

/* whole-batch section state [P][A][C][D] (mpi.cc:286-288 ships pop; Imap/Idiag are re-derived from it) */
extern "C" int beep_get_pop_all(BeepBatch *bt, double *out)
{
	if (!bt || !out) return fail(BEEP_ERR_INVALID, "null argument");
	const DevModel &m = bt->eng->dm;
	const size_t CD = (size_t)m.C * m.D;
	CU(cudaSetDevice(bt->eng->device));
	if (bt->db.popi) return get_pop_i32(bt, bt->db.popi, (size_t)bt->db.P * m.A, out);
	CU(cudaMemcpy2DAsync(out, CD * sizeof(double), bt->db.pop_cur, (size_t)m.CDp * sizeof(double), CD * sizeof(double), (size_t)bt->db.P * m.A, cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}
extern "C" int beep_set_pop_all(BeepBatch *bt, const double *in, int cur_t)
{
	if (!bt || !in) return fail(BEEP_ERR_INVALID, "null argument");
	const DevModel &m = bt->eng->dm;
	if (cur_t < 0 || cur_t > m.T) return fail(BEEP_ERR_INVALID, "step %d out of range", cur_t);
	const size_t CD = (size_t)m.C * m.D;
	CU(cudaSetDevice(bt->eng->device));
	CU(cudaStreamSynchronize(bt->stream));
	if (bt->db.popi) {
		std::vector<int> tmp((size_t)bt->db.P * m.A * CD);
		for (size_t i = 0; i < tmp.size(); i++) tmp[i] = (int)in[i];
		CU(cudaMemcpy2D(bt->db.popi, (size_t)m.CDp4 * sizeof(int), tmp.data(), CD * sizeof(int), CD * sizeof(int), (size_t)bt->db.P * m.A, cudaMemcpyHostToDevice));
		bt->cur_t = cur_t;
		return BEEP_OK;
	}
	CU(cudaMemcpy2D(bt->db.pop_cur, (size_t)m.CDp * sizeof(double), in, CD * sizeof(double), CD * sizeof(double), (size_t)bt->db.P * m.A, cudaMemcpyHostToDevice));
	bt->cur_t = cur_t;
	return BEEP_OK;
}

extern "C" int beep_get_derived(BeepBatch *bt, int slot, int which, double *out, int64_t n)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	if (!out) return fail(BEEP_ERR_INVALID, "null output");
	if (!bt->have_params) return fail(BEEP_ERR_STATE, "no parameters set");
	const DevModel &m = bt->eng->dm;
	const DevBatch &b = bt->db;
	CU(cudaSetDevice(bt->eng->device));
	if (bt->theta_dirty) { rc = launch_set_param(bt); if (rc) return rc; }
	const double *src = nullptr; size_t cnt = 0;
	switch (which) {
	case BEEP_DERIVED_TRANSRATE: cnt = (size_t)m.J * m.D; src = b.transrate + slot * cnt; break;
	case BEEP_DERIVED_DISC_SPLINE: cnt = (size_t)m.nspline * m.T; src = b.disc_spline + slot * cnt; break;
	case BEEP_DERIVED_SUSCEPTIBILITY: cnt = m.D; src = b.sus + slot * cnt; break;
	case BEEP_DERIVED_AREAFACTOR: cnt = (size_t)m.period * m.A; src = b.areafactor + slot * cnt; break;
	case BEEP_DERIVED_BETA: {
		/* expand the per-R-group series to the reference's [S][A][T] */
		cnt = (size_t)m.S * m.A * m.T;
		if ((size_t)n != cnt) return fail(BEEP_ERR_INVALID, "beta has %zu elements, not %lld", cnt, (long long)n);
		size_t gcnt = (size_t)m.S * m.nR * m.T;
		std::vector<double> g(gcnt);
		std::vector<int> rg(m.A);
		CU(cudaMemcpyAsync(g.data(), b.betaR + slot * gcnt, gcnt * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
		CU(cudaMemcpyAsync(rg.data(), m.area_rgrp, m.A * sizeof(int), cudaMemcpyDeviceToHost, bt->stream));
		CU(cudaStreamSynchronize(bt->stream));
		for (int st = 0; st < m.S; st++) for (int c = 0; c < m.A; c++)
			memcpy(out + ((size_t)st * m.A + c) * m.T, g.data() + ((size_t)st * m.nR + rg[c]) * m.T, m.T * sizeof(double));
		if (!bt->eng->h_mod_beta.empty())                              /* model.cc:1349-1357 */
			for (int st = 0; st < m.S; st++) for (int c = 0; c < m.A; c++) for (int t = 0; t < m.T; t++)
				out[((size_t)st * m.A + c) * m.T + t] *= bt->eng->h_mod_beta[((size_t)t * m.A + c) * m.S + st];
		return BEEP_OK;
	}
	case BEEP_DERIVED_AGEFACTOR: {
		/* g[t][a] = prod_mm spline_mm(t)^cnt_mm[a]; Ntime[t][a][b] = N0[a][b] g[t][a] g[t][b] up to rounding order */
		cnt = (size_t)m.T * m.G;
		if ((size_t)n != cnt) return fail(BEEP_ERR_INVALID, "agefactor has %zu elements, not %lld", cnt, (long long)n);
		size_t scnt = (size_t)m.nspline * m.T;
		std::vector<double> ds(scnt);
		std::vector<int> mmc((size_t)std::max(m.nmm, 1) * m.G), mms(std::max(m.nmm, 1));
		CU(cudaMemcpyAsync(ds.data(), b.disc_spline + slot * scnt, scnt * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
		if (m.nmm) {
			CU(cudaMemcpyAsync(mmc.data(), m.mm_cnt, (size_t)m.nmm * m.G * sizeof(int), cudaMemcpyDeviceToHost, bt->stream));
			CU(cudaMemcpyAsync(mms.data(), m.mm_spline, m.nmm * sizeof(int), cudaMemcpyDeviceToHost, bt->stream));
		}
		CU(cudaStreamSynchronize(bt->stream));
		for (int t = 0; t < m.T; t++) for (int a = 0; a < m.G; a++) {
			double x = 1;
			for (int mm = 0; mm < m.nmm; mm++) for (int k = 0; k < mmc[(size_t)mm * m.G + a]; k++) x *= ds[(size_t)mms[mm] * m.T + t];
			out[(size_t)t * m.G + a] = x;
		}
		return BEEP_OK;
	}
	default: return fail(BEEP_ERR_INVALID, "unknown derived selector %d", which);
	}
	if ((size_t)n != cnt) return fail(BEEP_ERR_INVALID, "selector %d has %zu elements, not %lld", which, cnt, (long long)n);
	CU(cudaMemcpyAsync(out, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
	CU(cudaStreamSynchronize(bt->stream));
	return BEEP_OK;
}

extern "C" int beep_export_particle(BeepBatch *bt, int slot, double *theta_host, double *transnum_host)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	const DevModel &m = bt->eng->dm;
	CU(cudaSetDevice(bt->eng->device));
	if (theta_host) {
		CU(cudaMemcpyAsync(theta_host, bt->db.theta + (size_t)slot * m.nparam, m.nparam * sizeof(double), cudaMemcpyDeviceToHost, bt->stream));
		CU(cudaStreamSynchronize(bt->stream));
	}
	if (transnum_host) return beep_get_trajectory(bt, slot, BEEP_TRAJ_TRANSNUM, transnum_host, (int64_t)m.T * m.A * m.J * m.D);
	return BEEP_OK;
}

extern "C" int beep_import_particle(BeepBatch *bt, int slot, const double *theta_host, const double *transnum_host)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	const DevModel &m = bt->eng->dm;
	CU(cudaSetDevice(bt->eng->device));
	CU(cudaStreamSynchronize(bt->stream));
	if (theta_host) {
		CU(cudaMemcpy(bt->db.theta + (size_t)slot * m.nparam, theta_host, m.nparam * sizeof(double), cudaMemcpyHostToDevice));
		bt->theta_dirty = true;
		/* the batch has parameters once every slot was given some (State::load per particle, state.cc:743-790) */
		if (bt->slot_theta.empty()) bt->slot_theta.assign(bt->db.P, 0);
		bt->slot_theta[slot] = 1;
		if (!bt->have_params) bt->have_params = std::find(bt->slot_theta.begin(), bt->slot_theta.end(), (unsigned char)0) == bt->slot_theta.end();
	}
	if (transnum_host) {
		const size_t TAJD = (size_t)m.T * m.A * m.J * m.D;
		if (m.stochastic) {
			if (!bt->db.tn) return fail(BEEP_ERR_STATE, "batch does not keep transnum");
			std::vector<int> tmp(TAJD);
			for (size_t i = 0; i < TAJD; i++) tmp[i] = (int)transnum_host[i];
			CU(cudaMemcpy(bt->db.tn + (size_t)slot * TAJD, tmp.data(), TAJD * sizeof(int), cudaMemcpyHostToDevice));
		} else {
			if (!bt->db.tm) return fail(BEEP_ERR_STATE, "batch does not keep transmean");
			CU(cudaMemcpy(bt->db.tm + (size_t)slot * TAJD, transnum_host, TAJD * sizeof(double), cudaMemcpyHostToDevice));
		}
	}
	/* the derived tables are re-derived from the stored parameters before the next launch (or read-back)
	 * that needs them: importing n particles costs one K0, not n */
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* PMCMC bootstrap, pmcmc.cc:211-268 (host, O(Ntot))                          */
/* ------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------ */
/* N4  posterior sample files, State::save / State::load (state.cc:709-790) */
/* ------------------------------------------------------------------------ */
/* Byte-compatible with the reference's Posterior/samples/sampleN_{parameter,transition}.csv: the same
 * std::ostream default formatting of doubles (6 significant digits), the same row and column order
 * [sett][area][trans][dp], the same mid-step times details.division_time (details.cc:102-105). */
extern "C" int beep_save_sample(BeepBatch *bt, int slot, const char *file_prefix, const char *const *param_names, double EF, double start_time)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	if (!file_prefix || !param_names) return fail(BEEP_ERR_INVALID, "null argument");
	const DevModel &m = bt->eng->dm;
	const size_t TAJD = (size_t)m.T * m.A * m.J * m.D, AJD = (size_t)m.A * m.J * m.D;
	std::vector<double> theta(m.nparam), tn(TAJD);
	rc = beep_export_particle(bt, slot, theta.data(), tn.data());
	if (rc) return rc;
	{
		std::string name = std::string(file_prefix) + "_parameter.csv";
		std::ofstream outp(name.c_str());
		if (!outp) return fail(BEEP_ERR_INVALID, "cannot open the file '%s'", name.c_str());
		outp << "Parameter,Value" << std::endl;
		outp << "Error function," << EF << std::endl;
		for (int th = 0; th < m.nparam; th++) outp << param_names[th] << "," << theta[th] << std::endl;
	}
	{
		std::string name = std::string(file_prefix) + "_transition.csv";
		std::ofstream outp(name.c_str());
		if (!outp) return fail(BEEP_ERR_INVALID, "cannot open the file '%s'", name.c_str());
		outp << "Time,Transitions (ordered by area x trans x democatpos)" << std::endl;
		const int period = m.period;
		for (int sett = 0; sett < m.T; sett++) {
			outp << (double((sett + 0.5) * period) / m.T + start_time);
			const double *row = tn.data() + (size_t)sett * AJD;
			for (size_t k = 0; k < AJD; k++) outp << "," << row[k];
			outp << std::endl;
		}
	}
	return BEEP_OK;
}

extern "C" int beep_load_sample(BeepBatch *bt, int slot, const char *file_prefix, const char *const *param_names, double *EF_out)
{
	int rc = slot_ok(bt, slot);
	if (rc) return rc;
	if (!file_prefix) return fail(BEEP_ERR_INVALID, "null argument");
	const DevModel &m = bt->eng->dm;
	const size_t AJD = (size_t)m.A * m.J * m.D;
	std::vector<double> theta(m.nparam), tn((size_t)m.T * AJD, 0.0);
	auto split = [](const std::string &ln) { std::vector<std::string> v; std::stringstream ss(ln); std::string f; while (std::getline(ss, f, ',')) v.push_back(f); return v; };
	{
		std::string name = std::string(file_prefix) + "_parameter.csv", line;
		std::ifstream inp(name.c_str());
		if (!inp) return fail(BEEP_ERR_INVALID, "cannot open the file '%s'", name.c_str());
		std::getline(inp, line);
		std::getline(inp, line);
		auto f = split(line);
		if (f.size() != 2) return fail(BEEP_ERR_INVALID, "problem loading the file '%s'", name.c_str());
		if (EF_out) *EF_out = atof(f[1].c_str());
		for (int th = 0; th < m.nparam; th++) {
			if (!std::getline(inp, line)) return fail(BEEP_ERR_INVALID, "the file '%s' is too short", name.c_str());
			f = split(line);
			if (f.size() != 2 || (param_names && f[0] != param_names[th])) return fail(BEEP_ERR_INVALID, "parameter %d of '%s' is not '%s'", th, name.c_str(), param_names ? param_names[th] : "?");
			theta[th] = atof(f[1].c_str());
		}
	}
	{
		std::string name = std::string(file_prefix) + "_transition.csv", line;
		std::ifstream inp(name.c_str());
		if (!inp) return fail(BEEP_ERR_INVALID, "cannot open the file '%s'", name.c_str());
		std::getline(inp, line);
		for (int sett = 0; sett < m.T; sett++) {
			if (!std::getline(inp, line)) break;                          /* a posterior shorter than the prediction period (simulate.cc:153) */
			auto f = split(line);
			if (f.size() != AJD + 1) return fail(BEEP_ERR_INVALID, "row %d of '%s' has %zu fields, not %zu", sett, name.c_str(), f.size(), AJD + 1);
			for (size_t k = 0; k < AJD; k++) tn[(size_t)sett * AJD + k] = atof(f[k + 1].c_str());
		}
	}
	return beep_import_particle(bt, slot, theta.data(), tn.data());
}

/* Mbp::get_al (mbp.cc:334-361) for n chains: the acceptance probability of a model-based proposal from the error functions
 * (EF = -2 log observation probability) and log priors of the initial and the proposed state.  mode 0 = CUTOFF (the ABC methods:
 * accept with the prior ratio when EF_p < EFcut), 1 = INVT (PAS, MC3, MCMC-MBP: power posterior at inverse temperature invT) */
extern "C" int beep_mbp_al(int mode, double invT, double EFcut, const double *EF_i, const double *EF_p, const double *Pr_i, const double *Pr_p, int n, double *al_out)
{
	if (!EF_i || !EF_p || !Pr_i || !Pr_p || !al_out || n < 0 || (mode != 0 && mode != 1)) return fail(BEEP_ERR_INVALID, "bad argument to beep_mbp_al");
	for (int i = 0; i < n; i++) {
		double al = 0.0;
		if (mode == 0) { if (EF_p[i] < EFcut) al = exp(Pr_p[i] - Pr_i[i]); }
		else al = exp(-invT * 0.5 * (EF_p[i] - EF_i[i]) + Pr_p[i] - Pr_i[i]);
		if (std::isnan(al)) return fail(BEEP_ERR_NUMERIC, "acceptance probability of chain %d is not a number (emsgEC Mbp 2)", i);
		if (al > 1) al = 1;
		al_out[i] = al;
	}
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* a12  resampling decisions of the particle methods (host arithmetic)      */
/* ------------------------------------------------------------------------ */
/* The reference gathers per-particle scalars on rank 0, decides there with the global ran() and
 * broadcasts; here every rank holds the all-gathered scalars and computes the identical decision from a
 * counter-based stream keyed by (seed, generation / sweep), so there is no root and no broadcast. */
static void decision_stream(BeepStream *rs, uint64_t seed, uint32_t key, uint32_t tag) { beep_stream_init(rs, seed, key, tag, 0xFFFFFFFEu); }

extern "C" int beep_test_uniforms(uint64_t seed, uint32_t key, uint32_t tag, int n, double *out)
{
	if (!out || n < 0) return fail(BEEP_ERR_INVALID, "bad argument to beep_test_uniforms");
	BeepStream rs;
	decision_stream(&rs, seed, key, tag);
	for (int i = 0; i < n; i++) out[i] = beep_next_uniform(&rs);
	return BEEP_OK;
}

/* ABCMBP::EF_cutoff, abcmbp.cc:135-217: cut-off = mean of the two middle error functions (:149-150); the
 * j-th best particle is kept and overwrites the (j + Ntot/2)-th (:158-163) */
extern "C" int beep_ef_cutoff(const double *EF, int ntot, double *EFcut, double *EFmin, double *EFmax, uint32_t *partcopy)
{
	if (!EF || ntot < 2 || !EFcut || !EFmin || !EFmax || !partcopy) return fail(BEEP_ERR_INVALID, "bad argument to beep_ef_cutoff");
	std::vector<uint32_t> ord(ntot);
	for (int i = 0; i < ntot; i++) ord[i] = (uint32_t)i;
	/* std::sort with the reference's comparator (abcmbp.cc:130-131,148): ties come out in the order the
	 * reference's own sort leaves them (tests/golden/decisions.npz holds a tie) */
	std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return EF[a] < EF[b]; });
	const int half = ntot / 2;
	*EFcut = 0.5 * (EF[ord[half - 1]] + EF[ord[half]]);
	*EFmin = EF[ord[(int)(0.025 * ntot)]]; *EFmax = EF[ord[(int)(0.975 * ntot)]];
	for (int i = 0; i < ntot; i++) partcopy[i] = 0xFFFFFFFFu;
	for (int j = 0; j < half; j++) partcopy[ord[j + half]] = ord[j];
	return BEEP_OK;
}

/* PAS::bootstrap, pas.cc:127-209 */
extern "C" int beep_pas_bootstrap(const double *EF, int nrun, int npart, double invT, double quench_factor, double invT_final,
                                  uint64_t seed, int gen, double *DinvT_out, double *EFmin, double *EFmax, uint32_t *partcopy)
{
	if (!EF || nrun <= 0 || npart <= 0 || !DinvT_out || !EFmin || !EFmax || !partcopy) return fail(BEEP_ERR_INVALID, "bad argument to beep_pas_bootstrap");
	const int ntot = nrun * npart;
	std::vector<double> srt(EF, EF + ntot);
	std::sort(srt.begin(), srt.end());
	*EFmin = srt[(int)(0.25 * ntot)]; *EFmax = srt[(int)(0.75 * ntot)];              /* :143 */
	const double EF_range = srt[(int)(0.5 * ntot)] - srt[(int)(0.25 * ntot)];         /* :147 */
	double muav = 0;
	for (int i = 0; i < ntot; i++) muav += EF[i];
	muav /= ntot;
	double DinvT = 0.25 * quench_factor / EF_range;                                   /* :154 */
	if (invT_final >= 0 && invT + DinvT > invT_final) DinvT = invT_final - invT;      /* :156-158 */
	*DinvT_out = DinvT;
	BeepStream rs;
	decision_stream(&rs, seed, (uint32_t)gen, 0x9A5u);
	std::vector<double> wsum(npart);
	std::vector<unsigned> num(npart);
	std::vector<uint32_t> list;
	for (int ru = 0; ru < nrun; ru++) {
		double sum = 0;
		for (int j = 0; j < npart; j++) { sum += exp(-0.5 * DinvT * (EF[ru * npart + j] - muav)); wsum[j] = sum; }
		for (int j = 0; j < npart; j++) num[j] = 0;
		for (int j = 0; j < npart; j++) {
			const double z = beep_next_uniform(&rs) * sum;
			int sel = 0;
			while (sel < npart - 1 && z > wsum[sel]) sel++;                              /* :170 (the last particle absorbs rounding) */
			num[sel]++;
		}
		list.clear();
		for (int j = 0; j < npart; j++) for (unsigned k = 1; k < num[j]; k++) list.push_back((uint32_t)(ru * npart + j));
		for (int j = 0; j < npart; j++) {
			if (num[j] > 0) partcopy[ru * npart + j] = 0xFFFFFFFFu;
			else { partcopy[ru * npart + j] = list.back(); list.pop_back(); }             /* :186-188: handed out from the end */
		}
	}
	return BEEP_OK;
}

/* MC3::swap_states, mc3.cc:153-217: five passes of neighbour swaps inside each run */
extern "C" int beep_mc3_swap(const uint32_t *run, const double *invT, const double *EF_in, int ntot, uint64_t seed, int sweep,
                             uint32_t *order, uint32_t *ntr_inc, uint32_t *nac_inc)
{
	if (!run || !invT || !EF_in || ntot <= 0 || !order || !ntr_inc || !nac_inc) return fail(BEEP_ERR_INVALID, "bad argument to beep_mc3_swap");
	std::vector<double> EF(EF_in, EF_in + ntot);
	for (int i = 0; i < ntot; i++) { order[i] = (uint32_t)i; ntr_inc[i] = 0; nac_inc[i] = 0; }
	BeepStream rs;
	decision_stream(&rs, seed, (uint32_t)sweep, 0x3C3u);
	for (int loop = 0; loop < 5; loop++)
		for (int i = 0; i + 1 < ntot; i++) {
			if (run[i] != run[i + 1]) continue;
			const double u = beep_next_uniform(&rs);                                     /* one per pair of the same run, as the reference's ran() (:180) */
			const double al = exp(0.5 * (invT[i] - invT[i + 1]) * (EF[i] - EF[i + 1]));   /* :179 */
			ntr_inc[i]++;
			if (u < al) {
				nac_inc[i]++;
				std::swap(order[i], order[i + 1]);
				std::swap(EF[i], EF[i + 1]);
			}
		}
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* K7  ABC-SMC importance weights                                            */
/* ------------------------------------------------------------------------ */
extern "C" int beep_mvn_kernel_sums(BeepEngine *eng, int nvar, int nparam, const int *var, const double *inv_M, double size,
                                    const double *param_prop, int N, const double *param_last, const double *w_last, int Nlast, double *sums_out)
{
	if (!eng || !var || !inv_M || !param_prop || !param_last || !w_last || !sums_out) return fail(BEEP_ERR_INVALID, "null argument");
	if (nvar <= 0 || nparam < nvar || N <= 0 || Nlast <= 0 || !(size > 0)) return fail(BEEP_ERR_INVALID, "bad sizes for beep_mvn_kernel_sums");
	for (int v = 0; v < nvar; v++) if (var[v] < 0 || var[v] >= nparam) return fail(BEEP_ERR_INVALID, "var[%d] = %d out of range", v, var[v]);
	/* inv_M / size^2 = Lm Lm' on the host in extended precision (nvar ~ 10^2): the symmetric part is
	 * what the quadratic form sees */
	std::vector<long double> L((size_t)nvar * nvar, 0.0L);
	for (int a = 0; a < nvar; a++)
		for (int b = 0; b <= a; b++) {
			long double s = 0.5L * ((long double)inv_M[(size_t)a * nvar + b] + (long double)inv_M[(size_t)b * nvar + a]) / ((long double)size * (long double)size);
			for (int k = 0; k < b; k++) s -= L[(size_t)a * nvar + k] * L[(size_t)b * nvar + k];
			if (a == b) {
				if (!(s > 0)) return fail(BEEP_ERR_INVALID, "inv_M is not positive definite (pivot %d)", a);
				L[(size_t)a * nvar + a] = sqrtl(s);
			} else L[(size_t)a * nvar + b] = s / L[(size_t)b * nvar + b];
		}
	std::vector<double> Lm((size_t)nvar * nvar), mu(nvar);
	for (size_t k = 0; k < Lm.size(); k++) Lm[k] = (double)L[k];
	{
		long double wsum = 0;
		std::vector<long double> acc(nvar, 0.0L);
		for (int j = 0; j < Nlast; j++) { wsum += w_last[j]; for (int v = 0; v < nvar; v++) acc[v] += (long double)w_last[j] * param_last[(size_t)j * nparam + var[v]]; }
		for (int v = 0; v < nvar; v++) mu[v] = (wsum > 0) ? (double)(acc[v] / wsum) : 0.0;
	}
	CU(cudaSetDevice(eng->device));
	cudaStream_t st = 0;
	const int ti = (N + MVN_TI - 1) / MVN_TI;
	int nslice = std::max(1, std::min((Nlast + MVN_TJ - 1) / MVN_TJ, (4 * eng->nsm + ti - 1) / ti));
	int jslice = (((Nlast + nslice - 1) / nslice) + MVN_TJ - 1) / MVN_TJ * MVN_TJ;
	nslice = (Nlast + jslice - 1) / jslice;
	/* one grow-only device scratch per engine, carved into 256-byte aligned pieces */
	auto al = [](size_t n) { return (n + 255) & ~(size_t)255; };
	const size_t sz[10] = {al((size_t)N * nparam * 8), al((size_t)Nlast * nparam * 8), al((size_t)Nlast * 8), al(Lm.size() * 8), al((size_t)nvar * 8),
	                       al((size_t)nvar * 4), al((size_t)N * nvar * 8), al((size_t)Nlast * nvar * 8), al((size_t)nslice * N * 8), al((size_t)N * 8)};
	size_t tot = 0;
	for (size_t x : sz) tot += x;
	if (tot > eng->mvn_cap) {
		if (eng->mvn_buf) { cudaFree(eng->mvn_buf); eng->mvn_buf = nullptr; eng->mvn_cap = 0; }
		CU(cudaMalloc(&eng->mvn_buf, tot));
		eng->mvn_cap = tot;
	}
	char *base = (char *)eng->mvn_buf;
	size_t off = 0;
	auto take = [&](int k) { char *p = base + off; off += sz[k]; return p; };
	double *d_prop = (double *)take(0), *d_last = (double *)take(1), *d_w = (double *)take(2), *d_L = (double *)take(3), *d_mu = (double *)take(4);
	int *d_var = (int *)take(5);
	double *d_xw = (double *)take(6), *d_yw = (double *)take(7), *d_part = (double *)take(8), *d_out = (double *)take(9);
#define MV(x) CU(x)
	MV(cudaMemcpyAsync(d_prop, param_prop, (size_t)N * nparam * sizeof(double), cudaMemcpyHostToDevice, st));
	MV(cudaMemcpyAsync(d_last, param_last, (size_t)Nlast * nparam * sizeof(double), cudaMemcpyHostToDevice, st));
	MV(cudaMemcpyAsync(d_w, w_last, (size_t)Nlast * sizeof(double), cudaMemcpyHostToDevice, st));
	MV(cudaMemcpyAsync(d_L, Lm.data(), Lm.size() * sizeof(double), cudaMemcpyHostToDevice, st));
	MV(cudaMemcpyAsync(d_mu, mu.data(), nvar * sizeof(double), cudaMemcpyHostToDevice, st));
	MV(cudaMemcpyAsync(d_var, var, nvar * sizeof(int), cudaMemcpyHostToDevice, st));
	k_mvn_whiten<<<N, 128, 0, st>>>(d_prop, N, nparam, d_var, nvar, d_mu, d_L, d_xw);
	k_mvn_whiten<<<Nlast, 128, 0, st>>>(d_last, Nlast, nparam, d_var, nvar, d_mu, d_L, d_yw);
	k_mvn_pair_sums<<<dim3(ti, nslice), 256, 0, st>>>(d_xw, N, d_yw, d_w, Nlast, nvar, jslice, d_part);
	k_mvn_reduce<<<(N + 127) / 128, 128, 0, st>>>(d_part, N, nslice, d_out);
	eng->launches += 4;
	MV(cudaGetLastError());
	MV(cudaMemcpyAsync(sums_out, d_out, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
	MV(cudaStreamSynchronize(st));
#undef MV
	return BEEP_OK;
}

extern "C" int beep_bootstrap(const double *L, int ntot, double invT, uint64_t seed, int section, double *loglik_inc, uint32_t *backpart)
{
	if (!L || ntot <= 0 || !loglik_inc || !backpart) return fail(BEEP_ERR_INVALID, "bad argument to beep_bootstrap");
	double Lmax = -1e300;
	for (int i = 0; i < ntot; i++) if (L[i] > Lmax) Lmax = L[i];
	std::vector<double> wsum(ntot);
	double sum = 0;
	for (int i = 0; i < ntot; i++) { sum += exp(invT * (L[i] - Lmax)); wsum[i] = sum; }
	*loglik_inc = invT * Lmax + log(sum / ntot);
	BeepStream rs;
	beep_stream_init(&rs, seed, (uint32_t)section, 0xB0075u, 0xFFFFFFFEu);
	std::vector<unsigned char> flag(ntot, 0);
	std::vector<uint32_t> extra;
	for (int i = 0; i < ntot; i++) backpart[i] = 0xFFFFFFFFu;
	for (int i = 0; i < ntot; i++) {
		double z = beep_next_uniform(&rs) * sum;
		int lo = 0, hi = ntot - 1;
		while (lo < hi) { int mid = (lo + hi) / 2; if (z <= wsum[mid]) hi = mid; else lo = mid + 1; }
		if (!flag[lo]) { flag[lo] = 1; backpart[lo] = (uint32_t)lo; }      /* stay in place if chosen, pmcmc.cc:250-258 */
		else extra.push_back((uint32_t)lo);
	}
	for (int i = 0; i < ntot; i++) if (backpart[i] == 0xFFFFFFFFu) { backpart[i] = extra.back(); extra.pop_back(); }   /* pmcmc.cc:256-260 */
	return BEEP_OK;
}

/* ------------------------------------------------------------------------ */
/* sampler parity entry points (used by tests only; device samplers)          */
/* ------------------------------------------------------------------------ */
extern "C" int beep_test_poisson(uint64_t seed, uint64_t stream_id, const double *lam_host, int n, int32_t *out_host)
{
	double *dl = nullptr; int *dout = nullptr;
	CU(cudaMalloc((void **)&dl, n * sizeof(double))); CU(cudaMalloc((void **)&dout, n * sizeof(int)));
	CU(cudaMemcpy(dl, lam_host, n * sizeof(double), cudaMemcpyHostToDevice));
	k_poisson_many<<<(n + 127) / 128, 128>>>(seed, stream_id, dl, n, dout);
	CU(cudaMemcpy(out_host, dout, n * sizeof(int), cudaMemcpyDeviceToHost));
	cudaFree(dl); cudaFree(dout);
	return BEEP_OK;
}
extern "C" int beep_test_binomial(uint64_t seed, uint64_t stream_id, const double *p_host, const int32_t *n_host, int n, int32_t *out_host)
{
	double *dp = nullptr; int *dn = nullptr, *dout = nullptr;
	CU(cudaMalloc((void **)&dp, n * sizeof(double))); CU(cudaMalloc((void **)&dn, n * sizeof(int))); CU(cudaMalloc((void **)&dout, n * sizeof(int)));
	CU(cudaMemcpy(dp, p_host, n * sizeof(double), cudaMemcpyHostToDevice)); CU(cudaMemcpy(dn, n_host, n * sizeof(int), cudaMemcpyHostToDevice));
	k_binomial_many<<<(n + 127) / 128, 128>>>(seed, stream_id, dp, dn, n, dout);
	CU(cudaMemcpy(out_host, dout, n * sizeof(int), cudaMemcpyDeviceToHost));
	cudaFree(dp); cudaFree(dn); cudaFree(dout);
	return BEEP_OK;
}

#include "comm.cuh"
