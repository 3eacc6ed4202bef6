/* imap_dense.cuh -- the area-mixing contraction as a batched FP64 tensor-core GEMM (DMMA).
 *
 *   k_imap_dmma    Imap[b][c'][g] = SUM_c Mt[c'][c] dinf[b][c][g]      reference src/state.cc:204-216
 *
 * The reference scatters along row c of the sparse matrix genQ.M for every (strain, area, age), one particle at a
 * time.  When M is DENSE (every area mixes with every other: nnz/A >= 64) and there is a batch of particles, the
 * same contraction for ALL particles of a step is one GEMM  C[A x N] = Mt[A x A] B[A x N],  N = P S G columns
 * (b = (particle, strain), g = age), and the A x A matrix is streamed ONCE per step for the whole batch -- the reuse
 * the per-tile gather lacks (A - 1 dependent L2 loads per lane and tile-step).  FP64 tensor cores:
 * mma.sync.aligned.m8n8k4.row.col.f64 (DMMA in SASS), CTA tile 64 x 32, four warps of 32 x 16, operands staged
 * in shared memory 16 columns of k at a time, the next stage prefetched into registers.  The k order of every output is fixed (stages ascending,
 * four products per DMMA), so the result does not depend on the launch; it differs from the sequential sum of the
 * gather by rounding only (tests: 1e-10 relative, measured ~1e-15).
 */
#ifndef BEEP_IMAP_DENSE_CUH
#define BEEP_IMAP_DENSE_CUH

#define IMD_TM 64
#define IMD_TN 32
#define IMD_KC 16

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, const double a, const double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

/* Mt [A][A] row-major (row = destination area c', column = source area c; zero diagonal: Idiag is separate).
 * dinf, imap: [NB][A][G] with NB = P*S batches.  Column j of the GEMM is (b, g) = (j / G, j % G).
 * CTA tile 64 x 32 (many CTAs: the matrices of this path are small), four warps of 32 x 16; the next stage's
 * operands travel from global memory into registers while the current stage is multiplied. */
__global__ void __launch_bounds__(128) k_imap_dmma(const double *__restrict__ Mt, const double *__restrict__ dinf, double *__restrict__ imap, int A, int G, int NB)
{
	__shared__ double As[IMD_TM][IMD_KC + 1];          /* [row][k] */
	__shared__ double Bs[IMD_KC][IMD_TN + 2];          /* [k][col] */
	const int N = NB * G;
	const int row0 = blockIdx.y * IMD_TM, col0 = blockIdx.x * IMD_TN;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wr = (warp >> 1) * 32, wc = (warp & 1) * 16;       /* the warp's 32 x 16 block of the CTA tile */
	const int fr = lane >> 2, fk = lane & 3;                       /* fragment coordinates: A row / B column = lane/4, k = lane%4 */
	double acc[4][2][2];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
	/* this thread's staging elements: 8 of the A tile (row ar + 8 q, column ak), 4 of the B tile (row bk + 4 q, column bc) */
	const int ak = tid & 15, ar = tid >> 4;
	const int bc = tid & 31, bk = tid >> 5;
	const int jb = col0 + bc;
	const bool bok = jb < N;
	const size_t boff = bok ? ((size_t)(jb / G) * A) * G + (jb % G) : 0;     /* + k * G */
	double ra[8], rb[4];
	auto fetch = [&](int k0) {
#pragma unroll
		for (int q = 0; q < 8; q++) { const int gr = row0 + ar + 8 * q, gk = k0 + ak; ra[q] = (gr < A && gk < A) ? Mt[(size_t)gr * A + gk] : 0.0; }
#pragma unroll
		for (int q = 0; q < 4; q++) { const int gk = k0 + bk + 4 * q; rb[q] = (bok && gk < A) ? dinf[boff + (size_t)gk * G] : 0.0; }
	};
	fetch(0);
	for (int k0 = 0; k0 < A; k0 += IMD_KC) {
		__syncthreads();
#pragma unroll
		for (int q = 0; q < 8; q++) As[ar + 8 * q][ak] = ra[q];
#pragma unroll
		for (int q = 0; q < 4; q++) Bs[bk + 4 * q][bc] = rb[q];
		__syncthreads();
		if (k0 + IMD_KC < A) fetch(k0 + IMD_KC);
#pragma unroll
		for (int ks = 0; ks < IMD_KC; ks += 4) {
			double af[4], bf[2];
#pragma unroll
			for (int i = 0; i < 4; i++) af[i] = As[wr + i * 8 + fr][ks + fk];
#pragma unroll
			for (int j = 0; j < 2; j++) bf[j] = Bs[ks + fk][wc + j * 8 + fr];
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 2; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
		}
	}
	/* C fragment: row = lane/4, columns 2*(lane%4) and +1 */
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 2; j++)
#pragma unroll
			for (int q = 0; q < 2; q++) {
				const int gr = row0 + wr + i * 8 + fr, jc = col0 + wc + j * 8 + 2 * fk + q;
				if (gr < A && jc < N) { const int b = jc / G, g = jc - b * G; imap[((size_t)b * A + gr) * G + g] = acc[i][j][q]; }
			}
}

/* dinf of the batch's current population (FP64 tiles), state.cc:223-264: the first step of a launch */
__global__ void k_dinf_init(DevModel m, DevBatch b, double *dinf_out)
{
	const int G = m.G, Dq = m.Dq, D = m.D, S = m.S, A = m.A;
	const long long n = (long long)b.P * S * A * G;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int a = (int)(i % G);
		const long long r = i / G;
		const int c = (int)(r % A);
		const int st = (int)((r / A) % S), p = (int)(r / ((long long)A * S));
		const double *pop = b.pop_cur + ((size_t)p * A + c) * m.CDp;
		const int *il = b.inflist + (size_t)p * (m.C + 1);
		const double *infco = b.infco + (size_t)p * m.C;
		double sum = 0;
		for (int k = 1; k <= il[0]; k++) {
			const int co = il[k];
			const double inf = infco[co];
			const double *pc = pop + co * D + st * Dq;
			if (m.age_fast) { for (int dp = a; dp < Dq; dp += G) sum += inf * pc[dp]; }
			else { for (int dp = 0; dp < Dq; dp++) if (m.dp_age[dp] == a) sum += inf * pc[dp]; }
		}
		dinf_out[i] = sum;
	}
}

#endif
