/* graphs.cuh -- the series behind the reference's Posterior/state files, straight from the device buffers
 * (row N4 of SURVEY.md 8f).
 *
 *   k_graph_state     ObservationModel::get_graph_state     reference src/obsmodel.cc:113-170
 *
 * A time-series graph plots, per group of graph_step steps, the events of its table's transitions (scaled to a rate
 * per unit time) and the populations of its table's compartments (averaged over the group), summed over the graph's
 * areas and demographic groups; a marginal graph plots its observations' values.  The reference walks one particle's
 * nested vectors on the host; here one warp per (particle, point) strides over the (step, area, group) cells of the
 * point -- the group index is innermost in transnum / pop, so lanes read unit-stride runs -- and reduces in a fixed
 * tree (results differ from the reference's sequential sum by rounding only; counts are integers).
 */
#ifndef BEEP_GRAPHS_CUH
#define BEEP_GRAPHS_CUH

#include "kernels.cuh"

struct DevGraphs {
	int ngraph, npoint, gstep;
	const int *type, *table, *fspline, *area_off, *area, *dp_off, *dp, *point_off, *point_obs;
	const int *pt_graph, *pt_first;           /* per output point: its graph; first step (time series) or observation (marginal) */
	const double *factor;
	const int *table_trans_off, *table_trans, *table_comp_off, *table_comp;
};

__global__ void __launch_bounds__(256) k_graph_state(const DevModel m, const DevBatch b, const DevGraphs g, double *out)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const int p = blockIdx.y;
	const int T = m.T, A = m.A, J = m.J, C = m.C, D = m.D;
	const double *ds = b.disc_spline + (size_t)p * m.nspline * T;
	for (int pt = blockIdx.x * nw + warp; pt < g.npoint; pt += gridDim.x * nw) {
		const int gr = g.pt_graph[pt];
		double sum = 0.0;
		if (g.type[gr] == 1) {
			if (lane == 0) sum = b.obs_val[(size_t)p * m.nobs + g.pt_first[pt]];
		} else {
			const int dt = g.table[gr], spl = g.fspline[gr], sett0 = g.pt_first[pt];
			const int na = g.area_off[gr + 1] - g.area_off[gr], nd = g.dp_off[gr + 1] - g.dp_off[gr];
			const double fac_trans = g.factor[gr] * (double)m.spu / g.gstep, fac_pop = g.factor[gr] * 1.0 / g.gstep;
			const int ncell = g.gstep * na * nd;
			for (int x = lane; x < ncell; x += 32) {
				const int step = x / (na * nd), r = x - step * (na * nd), ia = r / nd, id = r - ia * nd;
				const int sett2 = sett0 + step, c = g.area[g.area_off[gr] + ia], dp = g.dp[g.dp_off[gr] + id];
				const double sf = (spl < 0) ? 1.0 : ds[spl * T + sett2];
				for (int q = g.table_trans_off[dt]; q < g.table_trans_off[dt + 1]; q++) {
					const size_t e = ((((size_t)p * T + sett2) * A + c) * J + g.table_trans[q]) * D + dp;
					const double num = fac_trans * (b.tn ? (double)b.tn[e] : b.tm[e]);
					sum += (spl < 0) ? num : num * sf;
				}
				for (int q = g.table_comp_off[dt]; q < g.table_comp_off[dt + 1]; q++) {
					const double num = fac_pop * b.poptraj[((((size_t)p * T + sett2) * A + c) * C + g.table_comp[q]) * D + dp];
					sum += (spl < 0) ? num : num * sf;
				}
			}
		}
		for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
		if (lane == 0) out[(size_t)p * g.npoint + pt] = sum;
	}
}

#endif
