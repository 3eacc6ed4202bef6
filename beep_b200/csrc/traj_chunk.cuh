/* traj_chunk.cuh -- the grid-spanning trajectory kernel, second generation: ONE WARP PER CHUNK OF TILES.
 *
 *   k_traj_chunk<MODE, KEEP, NT>   State::simulate / Mbp::mbp / State::initialise_from_particle
 *                                  reference src/state.cc:596-638, src/mbp.cc:252-330, src/state.cc:294-322
 *
 * Same arithmetic as k_traj<MODE, SYNC_GRID, KEEP> with tile_step_fast (kernels.cuh) -- every transition
 * mean is the same sequence of IEEE operations, every cell draws from its own Philox stream -- hence the
 * same bits as the CPU oracle; what changes is how the work is laid on the machine.  k_traj gives a warp ONE
 * (particle, area) tile at a time: 20 of 32 lanes idle through the force of infection, nine passes over 260
 * cells of which ~40 can fire, draws in half-empty passes, a 1.9 KB FP64 tile streamed by 120 cp.async.
 * Here a warp takes a chunk of NT consecutive areas of one particle-step and works it PHASE BY PHASE:
 *
 *   load     the chunk's populations -- int32 in HBM, NT * C * D * 4 contiguous bytes -- arrive by ONE
 *            cp.async.bulk (an elected lane arms an mbarrier; no per-lane address arithmetic) while
 *   FOI      the force of infection is computed for all NT * G (tile, age) pairs, 32 per pass, every lane
 *            busy: sparse gather of the neighbours' infectivity (state.cc:204-216), then the G x G age
 *            contraction (state.cc:547-578) from shared memory;
 *   scan     per tile, 32 cells per pass from the packed cell table: cells that can fire are compacted
 *            (ballot/popc) with their mean into ONE list for the whole chunk (small means from the front,
 *            large ones from the back);
 *   draw     on the pooled list, 32 active cells per pass (full passes instead of one ragged pass per tile);
 *            events are applied at once to the int32 population in shared memory (red.shared) -- safe because
 *            a tile's means are all computed before the list is flushed -- and counted into integer
 *            accumulators of the observations that list the cell;
 *   close    observation sums, infectivity of the new population for all pairs, write-back of the tiles
 *            that changed (16-byte stores).
 * The list holds CAP >= J*D entries; when the next tile might not fit it is flushed first, so dense models
 * degrade to a flush per tile (the old behaviour) and sparse ones pool a whole chunk.
 *
 * Populations are integers in stochastic mode (the reference stores them as doubles, state.hh:28): int32 in
 * HBM halves the per-step traffic to the contract's 2 * C * 4 bytes per unit.
 *
 * The per-particle step scheduler is the one of k_traj<.., SYNC_GRID, ..> (kernels.cuh): a chunk is one grab.
 */
#ifndef BEEP_TRAJ_CHUNK_CUH
#define BEEP_TRAJ_CHUNK_CUH

#include "kernels.cuh"

#ifndef CHUNK_NWARP
#define CHUNK_NWARP 4
#endif

__device__ __forceinline__ void sts32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_shared_add(unsigned a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int4 lds128i(unsigned a) { int4 v; asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }

/* mbarrier + bulk asynchronous copy (one elected lane moves a contiguous block global -> shared) */
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned phase)
{
	unsigned ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
	return ok != 0;
}
/* orders this thread's generic-proxy accesses before subsequent async-proxy (bulk copy) accesses */
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

/* per-tile scratch of a chunk, 32 bytes */
#define TI_BEAF 0    /* double  beta * areafactor of the tile (state.cc:410-411)          */
#define TI_ET 8      /* double  external force of infection per individual (state.cc:423) */
#define TI_IE0G 16   /* int     efoi_ref * G: row of efoi_agedist                         */
#define TI_NZ 20     /* int     some age of the tile has a non-zero infectivity vector    */
#define TI_K0 24     /* int     first entry of the tile's (sett, area) observation list   */
#define TI_NENT 28   /* int     entries in that list                                      */



template <int MODE, bool KEEP, int NT>
__device__ __forceinline__ void chunk_step(const TrajArgs &a, const int p, const int t, const int c0, const int nt, const int lane,
                                           const unsigned wb, const unsigned ca, unsigned &phase)
{
	const DevModel &m = a.m;
	const DevBatch &b = a.b;
	const TrajSm &L = a.L;
	const ChunkSm &Ls = a.Ls;
	const int T = m.T, A = m.A, G = m.G;
	const int JD = m.J * G, CDp4 = m.CDp4;
	const unsigned pop_a = wb + Ls.pop, lm_a = wb + Ls.lmean, lc_a = wb + Ls.lcell, foi_a = wb + Ls.foi, acc_a = wb + Ls.acc, ti_a = wb + Ls.tinfo, bar_a = wb + Ls.bar;
	const unsigned ctab_a = ca + L.ctab * 8, n0_a = ca + L.n0 * 8;
	const size_t dsz = (size_t)b.P * A * G;
	const double *dinf_cur = b.dinf + (t & 1) * dsz + (size_t)p * A * G;
	const double *ds = b.disc_spline + (size_t)p * (m.nspline * T);
	const int npair = nt * G;
	const unsigned invG = a.invG;
	const int GsB = L.Gs * 8;

	/* 1a. state.cc:204-216, 547-553: the infectivity vector I of every (tile, age) pair */
	if ((G & 1) && lane < nt) sts64(lm_a + lane * GsB + G * 8, 0.0);      /* odd G: the age contraction reads pairs, one element past the vector */
	if (lane < nt) {
		/* per-tile scalars, one lane per tile: beta * area factor (state.cc:410-411), external force (:423) */
		const int c = c0 + lane;
		const AreaInfo ai = m.area_info[c];
		double be = b.betaR[((size_t)p * m.nR + ai.rg) * T + t];
		be *= b.areafactor[((size_t)p * m.period + t / m.spu) * A + c];
		const unsigned tia = ti_a + lane * 32;
		sts64(tia + TI_BEAF, be);
		sts64(tia + TI_ET, b.etab[((size_t)p * m.nefoi + ai.ie0) * T + t]);
		sts32(tia + TI_IE0G, (unsigned)(ai.ie0 * G));
		sts32(tia + TI_NZ, 0u);
	}
	__syncwarp();
	{
		const double d = ds[m.geo_spline * T + t], omd = 1 - d;
#pragma unroll 1
		for (int q0 = 0; q0 < npair; q0 += 32) {
			int q = q0 + lane;
			const bool valid = q < npair;
			if (!valid) q = npair - 1;                  /* idle lanes repeat the last pair: every load stays in range */
			const int i = (int)(((unsigned)q * invG) >> 16), ag = q - i * G, c = c0 + i;
			const AreaInfo ai = m.area_info[c];
			const char *dcl = (const char *)(dinf_cur + ag);
			const Nbr *nb = m.nbr + ai.j0;
			double Imap = 0;
			/* groups of eight sources (the rows of the matrix are padded to eight with zero weights on the area itself:
			 * adding +0 changes no bit): first the eight L2 requests from the byte offsets alone (eight value registers),
			 * then the weights (L1) and the sum in the reference's order -- one L2 round trip per group instead of one per
			 * unrolled quarter, without the 32 registers eight whole entries would hold */
			const double dg = ld_l2(dinf_cur + c * G + ag);
			const int nn8 = (ai.nn + 7) & ~7;
#pragma unroll 1
			for (int j0 = 0; j0 < nn8; j0 += 8) {
				double v[8];
#pragma unroll
				for (int u = 0; u < 8; u++) v[u] = ld_l2((const double *)(dcl + __ldg(&nb[j0 + u].off)));
#pragma unroll
				for (int u = 0; u < 8; u++) Imap += v[u] * __ldg(&nb[j0 + u].val);
			}
			const double Idiag = dg * ai.diag;
			const double Iv = (Idiag + Imap) * d + ai.fac * Idiag * omd;
			if (valid) {
				sts64(lm_a + i * GsB + ag * 8, Iv);
				if (Iv != 0) sts32(ti_a + i * 32 + TI_NZ, 1u);
			}
		}
	}
	__syncwarp();
	/* 1b. state.cc:555-578 NMI = max(0, Ntime I) and the force of infection per (tile, age), state.cc:406-432 */
	{
		const bool ntfly = m.nmm != 0;
#pragma unroll 1
		for (int q0 = 0; q0 < npair; q0 += 32) {
			int q = q0 + lane;
			const bool valid = q < npair;
			if (!valid) q = npair - 1;
			const int i = (int)(((unsigned)q * invG) >> 16), ag = q - i * G;
			const unsigned tia = ti_a + i * 32;
			const bool nz = lds32(tia + TI_NZ) != 0u;
			double nv = 0;
			/* a neighbourhood without infectious individuals has I = 0, hence NMI = 0 exactly */
			if (__any_sync(FULLMASK, nz)) {
				const unsigned row = n0_a + ag * GsB, ivr = lm_a + i * GsB;
				double sum = 0;
				if (!ntfly) {
					const int h = (G + 1) >> 1;                 /* pairs of ages: the padding of the rows (Gs - G elements) is not read */
#pragma unroll 5
					for (int k = 0; k < h; k++) { const double2 n2 = lds128(row + k * 16), v2 = lds128(ivr + k * 16); sum += n2.x * v2.x; sum += n2.y * v2.y; }
				} else {
#pragma unroll 1
					for (int aa = 0; aa < G; aa++) sum += ntime_elem(m, ds, t, ag, aa, lds64(row + aa * 8)) * lds64(ivr + aa * 8);
				}
				if (sum < 0) sum = 0;
				if (nz) nv = sum;
			}
			const double be = lds64(tia + TI_BEAF), et = lds64(tia + TI_ET);
			const double agd = m.efoi_agedist[(int)lds32(tia + TI_IE0G) + ag];
			if (valid) sts64(foi_a + q * 8, be * nv + et * agd);
		}
	}
	__syncwarp();

	/* 2. the chunk's populations have arrived (bulk copy issued by the caller before the FOI phase) */
	while (!mbar_try_wait(bar_a, phase)) { }
	phase ^= 1u;

	/* per-tile observation lists (obsmodel.cc:174-217 inverted); population observations read the
	 * population BEFORE this step's events (obsmodel.cc:99-105) */
	double *partial = b.obs_partial + (size_t)p * m.nslot;
	{
		int myk0 = 0, mynp = 0;
		if (lane < nt) {
			const int cell = t * A + c0 + lane;
			myk0 = m.cell_off[cell];
			mynp = m.cell_npop[cell];
			sts32(ti_a + lane * 32 + TI_K0, (unsigned)myk0);
			sts32(ti_a + lane * 32 + TI_NENT, (unsigned)(m.cell_off[cell + 1] - myk0));
		}
		unsigned pm = __ballot_sync(FULLMASK, mynp > 0);
		while (pm) {
			const int i = __ffs(pm) - 1; pm &= pm - 1;
			const int k0 = __shfl_sync(FULLMASK, myk0, i), np = __shfl_sync(FULLMASK, mynp, i);
			const unsigned pa = pop_a + i * (CDp4 * 4);
#pragma unroll 1
			for (int k = k0; k < k0 + np; k++) {
				const ObsEnt en = m.cell_ent[k];
				const int ob = en.ob;
				const unsigned long long dmask = m.obs_dmask[ob];
				const double factor = m.obs_factor[ob];
				double val = 0;
				if (lane < G && ((dmask >> lane) & 1ull)) {
					unsigned cm = m.obs_cmask[ob];
					while (cm) {
						const int co = __ffs(cm) - 1; cm &= cm - 1;
						val += factor * (double)(int)lds32(pa + (co * G + lane) * 4);
					}
				}
				for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(FULLMASK, val, o);
				if (lane == 0 && val != 0) atomicAdd(&partial[en.slot], val);
			}
		}
	}
	__syncwarp();

	/* 3. scan the tiles' cells into the chunk's list; flush (draw + apply) whenever the next tile might not fit */
	const bool keep_tn = KEEP && (b.flags & BEEP_KEEP_TRANSNUM) != 0, keep_tm = KEEP && (b.flags & BEEP_KEEP_TRANSMEAN) != 0;
	const unsigned lt = (1u << lane) - 1u;
	const double dt = m.dt;
	const bool apply = (t < T - 1);
	const size_t tbase0 = (KEEP || MODE != MODE_SIM) ? (((size_t)p * T + t) * A + c0) * (size_t)JD : 0;
	const double *crl = b.crate + (size_t)p * L.JD32 + lane;
	asm volatile("" : "+l"(crl));                       /* kept in registers: not re-derived from the launch arguments every pass */
	const unsigned long long cell00 = ((unsigned long long)t * A + c0) * (unsigned long long)JD;
	const unsigned long long sid = b.stream_id[p];
	const int top = Ls.cap - 1;
	unsigned dirty = 0;
	{
		int itile = 0;
#pragma unroll 1
		while (true) {
			unsigned cnt2 = 0;                                  /* nsmall | nbig << 16 */
#pragma unroll 1
			for (; itile < nt && (int)((cnt2 & 0xffffu) + (cnt2 >> 16)) + JD <= Ls.cap; itile++) {
				const unsigned pa = pop_a + itile * (CDp4 * 4), fa = foi_a + itile * (G * 8);
				const size_t tbase = tbase0 + (size_t)itile * JD;
				int simu = 0;
				if (MODE == MODE_MBP && a.simu_mask) simu = a.simu_mask[((size_t)p * T + t) * A + c0 + itile];
				/* the per-cell rate multipliers come through L1/L2 (the shared-memory carve-out leaves little L1): the
				 * next pass's value is requested while this pass is worked */
				double crn = __ldg(crl);
#pragma unroll 1
				for (int e0 = 0; e0 < JD; e0 += 32) {
					const int e = e0 + lane;
					const double cr = crn;
					crn = __ldg(crl + (unsigned)(e0 + 32));    /* the allocation is padded by one pass: the last request stays in range */
					const unsigned w = lds32(ctab_a + e * 4);          /* padded to a multiple of 32 with never-firing cells */
					const int popi = (int)lds32(pa + ((w & 0xfffu) << 2));
					if (MODE == MODE_REPLAY) {
						/* the events are given (state.cc:294-322); only the means are recomputed */
						int num = 0;
						if (e < JD) {
							double mean = dt * (double)popi * cr;
							if ((w >> 24) == 1u) mean *= lds64(fa + ((w >> 15) & 0x1f8u));
							if (!(popi > 0)) mean = 0;
							num = b.tn[tbase + e];
							if (keep_tm) b.tm[tbase + e] = mean;
						}
						const unsigned bal = __ballot_sync(FULLMASK, num != 0);
						if (num != 0) {
							const int pos = (int)(cnt2 & 0xffffu) + __popc(bal & lt);
							sts16(lc_a + pos * 2, (unsigned)((itile << 12) | e)); sts64(lm_a + pos * 8, (double)num);
						}
						cnt2 += __popc(bal);
						continue;
					}
					if (!__any_sync(FULLMASK, popi > 0 && (w >> 25) == 0u)) {          /* kind 2 (and the padding) never fires */
						if (KEEP && e < JD) {
							if (keep_tm) b.tm[tbase + e] = 0;
							if (keep_tn) b.tn[tbase + e] = 0;
						}
						continue;
					}
					double mean = dt * (double)popi * cr;
					if ((w >> 24) == 1u) mean *= lds64(fa + ((w >> 15) & 0x1f8u));
					if (!(popi > 0)) mean = 0;
					const bool act = (mean != 0);
					if (KEEP && e < JD) {
						if (keep_tm) b.tm[tbase + e] = mean;
						if (keep_tn && !act) b.tn[tbase + e] = 0;
					}
					const bool big = act && (mean >= BEEP_POISSON_INV_MAX) && !(MODE == MODE_MBP && !simu);
					const unsigned bal = __ballot_sync(FULLMASK, act), balb = __ballot_sync(FULLMASK, big);
					if (act) {
						const int pos = big ? top - (int)((cnt2 >> 16) + __popc(balb & lt)) : (int)((cnt2 & 0xffffu) + __popc(bal & ~balb & lt));
						sts16(lc_a + pos * 2, (unsigned)((itile << 12) | e)); sts64(lm_a + pos * 8, mean);
					}
					cnt2 += __popc(bal & ~balb) + (__popc(balb) << 16);
				}
			}
			const int nsmall = (int)(cnt2 & 0xffffu), nbig = (int)(cnt2 >> 16);
			__syncwarp();

			/* draws on the pooled list: one cell, one stream (state.cc:624-626 / mbp.cc:286-311) */
			auto draw = [&](const int cid, const double mean) -> int {
				const int i = cid >> 12, e = cid & 0xfff;
				BeepStream rs;
				beep_stream_init_cell(&rs, a.seed, sid, cell00 + (unsigned long long)(i * JD + e));
				rs.rk = a.rk;
				if (MODE == MODE_SIM) return beep_poisson(&rs, mean);
				const size_t g = tbase0 + (size_t)(i * JD + e);
				const int num_i = a.init_tn[g];
				if (a.simu_mask && a.simu_mask[((size_t)p * T + t) * A + c0 + i]) return beep_poisson(&rs, mean);
				const double rate_i = a.init_tm[g];
				if (mean > rate_i) return num_i + beep_poisson(&rs, mean - rate_i);
				if (mean == rate_i) return num_i;
				if (num_i == 0) return 0;
				return beep_binomial(&rs, mean / rate_i, num_i);
			};
			/* one settled cell: trajectory store, population update (state.cc:460-485; the last step's events are
			 * never applied) and the integer event sums of the observations that list the cell (obsmodel.cc:81-97) */
			auto finish = [&](const int cid, const int num) {
				const int i = cid >> 12, e = cid & 0xfff;
				if (MODE != MODE_REPLAY && keep_tn) b.tn[tbase0 + (size_t)(i * JD + e)] = num;
				if (num != 0) {
					const unsigned w = lds32(ctab_a + e * 4);
					const int tr = (w >> 12) & 63, dp = (w >> 18) & 63;
					if (apply) {
						const unsigned pa = pop_a + i * (CDp4 * 4);
						red_shared_add(pa + (a.tfrom[tr] * G + dp) * 4, -num);
						red_shared_add(pa + (a.tto[tr] * G + dp) * 4, num);
						dirty |= 1u << i;
					}
					if ((m.obs_tr_any >> tr) & 1u) {
						const int k0 = (int)lds32(ti_a + i * 32 + TI_K0), nent = (int)lds32(ti_a + i * 32 + TI_NENT);
#pragma unroll 1
						for (int j = 0; j < nent; j++) {
							const int ob = m.cell_ent[k0 + j].ob;
							if (((m.obs_tmask[ob] >> tr) & 1u) && ((m.obs_dmask[ob] >> dp) & 1ull)) red_shared_add(acc_a + (i * CHUNK_NEJ + j) * 4, num);
						}
					}
				}
			};
			if (MODE == MODE_REPLAY) {
#pragma unroll 1
				for (int x = lane; x < nsmall; x += 32) finish((int)lds16(lc_a + x * 2), (int)lds64(lm_a + x * 8));
			} else {
#pragma unroll 1
				for (int x = lane; x < nsmall; x += 32) {
					const int cid = (int)lds16(lc_a + x * 2);
					finish(cid, draw(cid, lds64(lm_a + x * 8)));
				}
				if (nbig > 0) {
					/* MODE_SIM only (MBP keeps a single list): first the squeeze of every cell's first attempt; the cells
					 * it does not settle are collected (over the slots already read) and run the full sampler together */
					int ndef = 0;
#pragma unroll 1
					for (int x0 = 0; x0 < nbig; x0 += 32) {
						const int x = x0 + lane;
						bool defer = false;
						int cid = 0;
						double mean = 0;
						if (x < nbig) {
							cid = (int)lds16(lc_a + (top - x) * 2);
							mean = lds64(lm_a + (top - x) * 8);
							BeepStream rs;
							beep_stream_init_cell(&rs, a.seed, sid, cell00 + (unsigned long long)((cid >> 12) * JD + (cid & 0xfff)));
							int num;
							if (beep_poisson_ptrs_try(rs.key[0], rs.key[1], rs.ctr[0], rs.ctr[1], rs.ctr[2], rs.ctr[3], mean, &num, a.rk)) finish(cid, num);
							else defer = true;
						}
						const unsigned bd = __ballot_sync(FULLMASK, defer);
						if (defer) { const int pos = top - (ndef + __popc(bd & lt)); sts16(lc_a + pos * 2, (unsigned)cid); sts64(lm_a + pos * 8, mean); }
						ndef += __popc(bd);
					}
					__syncwarp();
#pragma unroll 1
					for (int x = lane; x < ndef; x += 32) {
						const int cid = (int)lds16(lc_a + (top - x) * 2);
						finish(cid, draw(cid, lds64(lm_a + (top - x) * 8)));
					}
				}
			}
			__syncwarp();
			if (itile >= nt) break;
		}
	}

	/* 4a. obsmodel.cc:74-97: the transition observations of the chunk's (sett, area) cells */
#pragma unroll 1
	for (int x = lane; x < nt * CHUNK_NEJ; x += 32) {
		const int v = (int)lds32(acc_a + x * 4);
		if (v != 0) {
			sts32(acc_a + x * 4, 0u);
			const int i = x / CHUNK_NEJ, j = x - i * CHUNK_NEJ;
			const ObsEnt en = m.cell_ent[(int)lds32(ti_a + i * 32 + TI_K0) + j];
			const double factor = m.obs_factor[en.ob];
			const int fs = m.obs_fs[en.ob];
			/* the slot has a single owner (this tile) and steps are ordered: a plain RED */
			atomicAdd(&partial[en.slot], (fs < 0) ? factor * (double)v : factor * (double)v * ds[fs * T + t]);
		}
	}
	/* 4b. state.cc:223-264: infectivity of the new population, all (tile, age) pairs; write back what changed */
	if (apply) {
		dirty = __reduce_or_sync(FULLMASK, dirty);
		double *dinf_nxt = b.dinf + ((t + 1) & 1) * dsz + (size_t)p * A * G + (size_t)c0 * G;
		const int *il = b.inflist + (size_t)p * (m.C + 1);
		const int ninf = il[0];
		const unsigned lu_a = wb + Ls.lu;                   /* free here: (coefficient, byte offset) of the infectious compartments */
		if (lane < ninf) { const int co = il[1 + lane]; sts64(lu_a + lane * 8, b.infco[(size_t)p * m.C + co]); sts32(lu_a + 64 + lane * 4, (unsigned)(co * G * 4)); }
		__syncwarp();
#pragma unroll 1
		for (int q = lane; q < npair; q += 32) {
			const int i = (int)(((unsigned)q * invG) >> 16), ag = q - i * G;
			const unsigned pa = pop_a + i * (CDp4 * 4) + ag * 4;
			double sum = 0;
#pragma unroll 2
			for (int k = 0; k < ninf; k++) sum += lds64(lu_a + k * 8) * (double)(int)lds32(pa + lds32(lu_a + 64 + k * 4));
			dinf_nxt[q] = sum;
		}
		const int n16 = CDp4 >> 2;                          /* 16-byte pieces per tile */
		int4 *gp = (int4 *)(b.popi + ((size_t)p * A + c0) * CDp4);
		while (dirty) {
			const int i = __ffs(dirty) - 1; dirty &= dirty - 1;
#pragma unroll 1
			for (int x = lane; x < n16; x += 32) gp[i * n16 + x] = lds128i(pop_a + (i * n16 + x) * 16);
		}
	}
	__syncwarp();
}

template <int NT> struct ChunkCfg { static constexpr int MINB = ((NT <= 4) ? 24 : 16) / CHUNK_NWARP; };

template <int MODE, bool KEEP, int NT>
__global__ void __launch_bounds__(CHUNK_NWARP * 32, ChunkCfg<NT>::MINB) k_traj_chunk(const __grid_constant__ TrajArgs a)
{
	extern __shared__ __align__(16) double smraw[];
	const DevModel &m = a.m;
	const DevBatch &b = a.b;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int A = m.A, G = m.G, CDp4 = m.CDp4;
	const TrajSm &L = a.L;
	const ChunkSm &Ls = a.Ls;
	{
		TileSm ts0;
		sm_common_init(m, L, smraw, ts0);
	}
	const unsigned ca = launder(smem_addr(smraw));
	const unsigned wb = launder(ca + L.common * 8 + warp * Ls.per);
	for (int x = lane; x < NT * CHUNK_NEJ; x += 32) sts32(wb + Ls.acc + x * 4, 0u);
	if (lane == 0) { mbar_init(wb + Ls.bar, 1u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); fence_proxy_async(); }
	__syncthreads();
	unsigned phase = 0;

	cg::grid_group grid = cg::this_grid();
	const int ntile = b.P * A, nw = gridDim.x * CHUNK_NWARP, gw = blockIdx.x * CHUNK_NWARP + warp;   /* 32-bit: the launcher checks P*A < 2^30 */
	const size_t dsz = (size_t)b.P * A * G;
	/* infectivity of the starting population (static tile split: done once) */
#pragma unroll 1
	for (int tile = gw; tile < ntile; tile += nw) {
		const int p = tile / A;
		const int *gp = b.popi + (size_t)tile * CDp4;
		const int *il = b.inflist + (size_t)p * (m.C + 1);
		const double *infco = b.infco + (size_t)p * m.C;
		if (lane < G) {
			const int ninf = il[0];
			double sum = 0;
#pragma unroll 1
			for (int k = 1; k <= ninf; k++) { const int co = il[k]; sum += infco[co] * (double)gp[co * G + lane]; }
			b.dinf[(a.ti & 1) * dsz + (size_t)tile * G + lane] = sum;
		}
	}
	grid.sync();

	/* the particle-step scheduler of k_traj<.., SYNC_GRID, ..> (kernels.cuh), one chunk of NT tiles per grab */
	const int P = b.P;
	int p = (int)(((long long)gw * P) / nw);
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
	int nmiss = 0, nfin = 0, nspin = 0;
	unsigned long long v = 0;
	bool have = false;
#pragma unroll 1
	while (true) {
		unsigned long long *sp = b.sched + (size_t)p * 16;
		if (!have) {
			if (lane == 0) {
				v = ld_relaxed_u64(sp);
				if ((int)(v >> 32) < a.tf && (unsigned)v < (unsigned)A) v = atomicAdd(sp, (unsigned long long)NT);
			}
		}
		have = false;
		v = __shfl_sync(FULLMASK, v, 0);
		const int t = (int)(v >> 32);
		const unsigned idx = (unsigned)v;
		if (t >= a.tf || idx >= (unsigned)A) {
			nmiss++;
			if (t >= a.tf) nfin++;
			/* every particle has closed its last step (counted by the closers): nothing is left anywhere */
			if ((nmiss & 3) == 0 && ld_relaxed_u64(b.sched + (size_t)P * 16) >= (unsigned long long)P) break;
			if (nmiss >= P) {
				if (nfin >= P) break;
				nmiss = 0; nfin = 0;
				if (++nspin > 8000000) __trap();
				__nanosleep(500);
			}
			p += hop;
			if (p >= P) p -= P;
			continue;
		}
		nmiss = 0; nfin = 0; nspin = 0;
		const int c0 = (int)idx, c1 = (c0 + NT < A) ? c0 + NT : A;
		/* the chunk's populations: one bulk copy, awaited after the force-of-infection phase.  The previous
		 * chunk's generic accesses to the buffer are ordered before it (proxy fence by every lane, warp barrier) */
		fence_proxy_async();
		__syncwarp();
		if (lane == 0) {
			const unsigned bytes = (unsigned)((c1 - c0) * CDp4 * 4);
			mbar_expect_tx(wb + Ls.bar, bytes);
			bulk_g2s(wb + Ls.pop, b.popi + ((size_t)p * A + c0) * CDp4, bytes, wb + Ls.bar);
		}
		unsigned long long vn = 0;
		if (c1 < A) { have = true; if (lane == 0) vn = atomicAdd(sp, (unsigned long long)NT); }
		chunk_step<MODE, KEEP, NT>(a, p, t, c0, c1 - c0, lane, wb, ca, phase);
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

/* int32 populations of a batch: reset to the initial state (state.cc:600) */
__global__ void k_pop_reset_i32(DevModel m, DevBatch b)
{
	const size_t CD = (size_t)m.C * m.D, CDp4 = m.CDp4, n = (size_t)m.A * CDp4, tot = n * b.P;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
		size_t r = i % n, c = r / CDp4, k = r - c * CDp4;
		b.popi[i] = (k < CD) ? (int)m.pop_init[c * CD + k] : 0;
	}
}

#endif
