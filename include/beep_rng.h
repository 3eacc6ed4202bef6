/* beep_rng.h -- counter-based random streams and exact Poisson / binomial
 * samplers with ONE floating-point op sequence on host and device.
 *
 * Replaces (reference src/utils.cc:32-34,148-153,174-180):
 *     std::default_random_engine generator;            // minstd_rand0, global
 *     poisson_sample(lam)   -> std::poisson_distribution<int>
 *     binomial_sample(p,n)  -> std::binomial_distribution<int>
 * whose algorithms are implementation-defined (libstdc++) and whose draw
 * order is a global sequential side effect.  Here every draw is a pure
 * function of a 64-bit key and a 128-bit counter (Philox4x32-10, Salmon et
 * al. SC'11), and the samplers use only IEEE-754 correctly rounded
 * operations (+ - * / sqrt, integer ops) written without contraction, plus
 * software log/exp built from those, so that the CUDA kernels, the CPU
 * oracle and the reference harness (which links this header in place of the
 * std:: distributions) produce bit-identical integers from the same stream.
 *
 * Samplers (both exact in distribution):
 *     Poisson(lam):  lam < 10  sequential-search inversion
 *                    lam >= 10 PTRS transformed rejection (Hormann 1993)
 *     Binomial(n,p): n*min(p,1-p) < 10  BINV inversion (Kachitvichyanukul
 *                    & Schmeiser 1988), else BTRS (Hormann 1993)
 *
 * Host compilation of any TU including this header must use
 * -ffp-contract=off (the build recipes do); device code uses the *_rn
 * intrinsics, which nvcc never contracts into FMAs.
 */
#ifndef BEEP_RNG_H
#define BEEP_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define BEEP_HD __host__ __device__ __forceinline__
/* rarely taken acceptance tests: kept out of line so that the hot loop of a kernel stays small */
#define BEEP_HD_COLD static __host__ __device__ __noinline__
#else
#define BEEP_HD static inline
#define BEEP_HD_COLD static
#include <math.h>
#endif

#if defined(__CUDA_ARCH__)
#define BR_MUL(a, b) __dmul_rn((a), (b))
#define BR_ADD(a, b) __dadd_rn((a), (b))
#define BR_SUB(a, b) __dsub_rn((a), (b))
#define BR_DIV(a, b) __ddiv_rn((a), (b))
#define BR_SQRT(a) __dsqrt_rn((a))
#define BR_MULHI(a, b) __umulhi((a), (b))
#define BR_FLOOR(a) floor((a))
#define BR_FABS(a) fabs((a))
#else
#define BR_MUL(a, b) ((a) * (b))
#define BR_ADD(a, b) ((a) + (b))
#define BR_SUB(a, b) ((a) - (b))
#define BR_DIV(a, b) ((a) / (b))
#define BR_SQRT(a) sqrt((a))
#define BR_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#define BR_FLOOR(a) floor((a))
#define BR_FABS(a) fabs((a))
#endif

/* ------------------------------------------------------------------------ */
/* Philox4x32-10                                                             */
/* ------------------------------------------------------------------------ */
typedef struct BeepStream {
	uint32_t key[2]; /* seed                                                   */
	uint32_t ctr[4]; /* ctr[0] = block index within the stream; ctr[1..3] = id */
	uint32_t buf[4];
	int have;        /* unread words in buf                                    */
	const uint32_t *rk; /* device kernels: the ten round keys (key + r * Weyl constants) precomputed by the launcher, or null */
} BeepStream;

BEEP_HD void beep_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int r = 0; r < 10; r++) {
		uint32_t hi0 = BR_MULHI(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		uint32_t hi1 = BR_MULHI(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#if defined(__CUDACC__)
/* The same block with the round keys given: rk[2r] = key[0] + r * 0x9E3779B9, rk[2r+1] = key[1] + r * 0xBB67AE85 (the launcher
 * computes them once per seed).  In a kernel they sit in the constant bank of the launch arguments and feed the three-input
 * XORs directly: four issue slots per round instead of six.  Same bits. */
static __device__ __forceinline__ void beep_philox4x32_10_rk(const uint32_t ctr[4], const uint32_t *rk, uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif
/* a block of the stream: through the precomputed round keys where the caller gave them */
BEEP_HD void beep_stream_block(BeepStream *s)
{
#if defined(__CUDA_ARCH__)
	if (s->rk) { beep_philox4x32_10_rk(s->ctr, s->rk, s->buf); return; }
#endif
	beep_philox4x32_10(s->ctr, s->key, s->buf);
}

/* A stream is identified by a 64-bit seed (the Philox key) and the upper three
 * counter words; ctr[0] counts the blocks consumed.  Two keyings are used:
 *   sequential  (id0, id1, 0xFFFFFFFF): ONE stream consumed in the reference's
 *               loop order -- reference harness and the oracle's SEQ mode;
 *   per cell    beep_stream_init_cell(): one stream per particle stream id and
 *               per (sett, area, transition, dp) cell -- the CUDA engine and
 *               the oracle's KEYED mode.  cell < 2^48; a draw uses < 2^16 blocks. */
BEEP_HD void beep_stream_init(BeepStream *s, uint64_t seed, uint32_t id0, uint32_t id1, uint32_t id2)
{
	s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
	s->ctr[0] = 0; s->ctr[1] = id0; s->ctr[2] = id1; s->ctr[3] = id2;
	s->have = 0; s->rk = 0;
}

BEEP_HD void beep_stream_init_cell(BeepStream *s, uint64_t seed, uint64_t stream_id, uint64_t cell)
{
	s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
	s->ctr[0] = (uint32_t)(cell >> 32) << 16; s->ctr[1] = (uint32_t)cell;
	s->ctr[2] = (uint32_t)stream_id; s->ctr[3] = (uint32_t)(stream_id >> 32);
	s->have = 0; s->rk = 0;
}

BEEP_HD uint32_t beep_next_u32(BeepStream *s)
{
	if (s->have == 0) {
		beep_stream_block(s);
		s->ctr[0]++;
		s->have = 4;
	}
	uint32_t w = s->buf[4 - s->have];
	s->have--;
	return w;
}

/* uniform on (0,1): (x + 1/2) 2^-52 with x a 52-bit integer; exact, never 0/1.
 * Always consumes an aligned pair of words, so a Philox block yields exactly
 * two uniforms (static buffer indices: the stream stays in registers on the GPU) */
BEEP_HD double beep_next_uniform(BeepStream *s)
{
	uint32_t a, b;
	if (s->have < 2) {
		beep_stream_block(s);
		s->ctr[0]++;
		s->have = 4;
	}
	if (s->have >= 4) { a = s->buf[0]; b = s->buf[1]; s->have = 2; }
	else { a = s->buf[2]; b = s->buf[3]; s->have = 0; }
	uint64_t x = (((uint64_t)a << 32) | (uint64_t)b) >> 12;
	return BR_MUL(BR_ADD((double)x, 0.5), 2.220446049250313080847263336181640625e-16);
}

/* ------------------------------------------------------------------------ */
/* software log / exp (fdlibm e_log.c / e_exp.c formulation, written with    */
/* explicit non-contracted operations; ~1 ulp, identical on host and device) */
/* ------------------------------------------------------------------------ */
BEEP_HD uint64_t beep_d2u(double x)
{
#if defined(__CUDA_ARCH__)
	return (uint64_t)__double_as_longlong(x);
#else
	union { double d; uint64_t u; } v; v.d = x; return v.u;
#endif
}
BEEP_HD double beep_u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long)u);
#else
	union { double d; uint64_t u; } v; v.u = u; return v.d;
#endif
}

/* natural log of a positive finite double (out of line on the device: only the rarely
 * taken exact acceptance tests of the rejection samplers call it) */
BEEP_HD_COLD double beep_log(double x)
{
	const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
	const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
	             Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
	             Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
	             Lg7 = 1.479819860511658591e-01;
	uint64_t ux = beep_d2u(x);
	int k = 0;
	if ((ux >> 52) == 0) { /* subnormal: scale by 2^54 */
		x = BR_MUL(x, 18014398509481984.0);
		ux = beep_d2u(x);
		k = -54;
	}
	int e = (int)((ux >> 52) & 0x7ff) - 1023;
	uint64_t m = ux & 0x000fffffffffffffULL;
	/* normalise to [sqrt(2)/2, sqrt(2)) */
	if (m >= 0x6a09e667f3bcdULL) { e += 1; ux = m | 0x3fe0000000000000ULL; }
	else ux = m | 0x3ff0000000000000ULL;
	k += e;
	double f = BR_SUB(beep_u2d(ux), 1.0);
	double dk = (double)k;
	double s = BR_DIV(f, BR_ADD(2.0, f));
	double z = BR_MUL(s, s);
	double w = BR_MUL(z, z);
	double t1 = BR_MUL(w, BR_ADD(Lg2, BR_MUL(w, BR_ADD(Lg4, BR_MUL(w, Lg6)))));
	double t2 = BR_MUL(z, BR_ADD(Lg1, BR_MUL(w, BR_ADD(Lg3, BR_MUL(w, BR_ADD(Lg5, BR_MUL(w, Lg7)))))));
	double R = BR_ADD(t2, t1);
	double hfsq = BR_MUL(0.5, BR_MUL(f, f));
	/* k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f) */
	double inner = BR_ADD(BR_MUL(s, BR_ADD(hfsq, R)), BR_MUL(dk, ln2_lo));
	return BR_SUB(BR_MUL(dk, ln2_hi), BR_SUB(BR_SUB(hfsq, inner), f));
}

/* exp(x) for x in [-700, 700] (callers guarantee the range).  Division-free:
 * x = k ln2 + r, |r| <= ln2/2 (two-part ln2 as in fdlibm), then
 * exp(r) = 1 + (r + r^2 q(r)) with q the degree-11 Taylor tail 1/2 + r/6 + ... + r^11/13!
 * (truncation < 5e-18), evaluated as two interleaved Horner chains in r^2 (even and odd
 * coefficients) with explicit non-contracted operations: < 1 ulp, identical on host and device. */
BEEP_HD double beep_exp(double x)
{
	const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
	             invln2 = 1.44269504088896338700e+00;
	const double c2 = 5.0e-01, c3 = 1.66666666666666666667e-01, c4 = 4.16666666666666666667e-02,
	             c5 = 8.33333333333333333333e-03, c6 = 1.38888888888888888889e-03, c7 = 1.98412698412698412698e-04,
	             c8 = 2.48015873015873015873e-05, c9 = 2.75573192239858906526e-06, c10 = 2.75573192239858906526e-07,
	             c11 = 2.50521083854417187751e-08, c12 = 2.08767569878680989792e-09, c13 = 1.60590438368216145994e-10;
	int k = (int)BR_ADD(BR_MUL(invln2, x), (x < 0.0 ? -0.5 : 0.5));
	double dk = (double)k;
	double r = BR_SUB(BR_SUB(x, BR_MUL(dk, ln2HI)), BR_MUL(dk, ln2LO));
	double z = BR_MUL(r, r);
	/* q(r) = E(z) + r O(z):  E = c2 + c4 z + ... + c12 z^5,  O = c3 + c5 z + ... + c13 z^5 */
	double E = BR_ADD(c2, BR_MUL(z, BR_ADD(c4, BR_MUL(z, BR_ADD(c6, BR_MUL(z, BR_ADD(c8, BR_MUL(z, BR_ADD(c10, BR_MUL(z, c12))))))))));
	double O = BR_ADD(c3, BR_MUL(z, BR_ADD(c5, BR_MUL(z, BR_ADD(c7, BR_MUL(z, BR_ADD(c9, BR_MUL(z, BR_ADD(c11, BR_MUL(z, c13))))))))));
	double q = BR_ADD(E, BR_MUL(r, O));
	double y = BR_ADD(1.0, BR_ADD(r, BR_MUL(z, q)));
	/* scale by 2^k, k in [-1010, 1010]: two-step to stay in the normal range */
	int k1 = k / 2, k2 = k - k1;
	double s1 = beep_u2d((uint64_t)(1023 + k1) << 52);
	double s2 = beep_u2d((uint64_t)(1023 + k2) << 52);
	return BR_MUL(BR_MUL(y, s1), s2);
}

/* log(k!) for integer k >= 0: table below 16, Stirling series above
 * (truncation error < 2e-14 at k = 16)                                     */
BEEP_HD double beep_lfact(double k)
{
	if (k < 16.0) {
		switch ((int)k) {
		case 0: case 1: return 0.0;
		case 2: return 0.693147180559945309417232121458;
		case 3: return 1.791759469228055000812477358381;
		case 4: return 3.178053830347945619646941601297;
		case 5: return 4.787491742782045994247700934523;
		case 6: return 6.579251212010100995060178292904;
		case 7: return 8.525161361065414300165531036347;
		case 8: return 10.60460290274525022841722740072;
		case 9: return 12.80182748008146961120771787457;
		case 10: return 15.10441257307551529522570932925;
		case 11: return 17.50230784587388583928765290722;
		case 12: return 19.98721449566188614951736238706;
		case 13: return 22.55216385312342288557084982862;
		case 14: return 25.19122118273868150009343469352;
		default: return 27.89927138384089156608943926367;
		}
	}
	/* lgamma(k+1) = (k+1/2) log k - k + log(2 pi)/2 + 1/(12k) - 1/(360k^3) + 1/(1260k^5) - 1/(1680k^7) */
	double ik = BR_DIV(1.0, k);
	double ik2 = BR_MUL(ik, ik);
	double ser = BR_MUL(ik, BR_ADD(8.333333333333333333e-02,
	             BR_MUL(ik2, BR_ADD(-2.777777777777777778e-03,
	             BR_MUL(ik2, BR_ADD(7.936507936507936508e-04,
	             BR_MUL(ik2, -5.952380952380952381e-04)))))));
	return BR_ADD(BR_ADD(BR_SUB(BR_MUL(BR_ADD(k, 0.5), beep_log(k)), k), 9.189385332046727417803297364056e-01), ser);
}

/* ------------------------------------------------------------------------ */
/* Poisson                                                                   */
/* ------------------------------------------------------------------------ */
#define BEEP_POISSON_INV_MAX 10.0

/* 1/k, correctly rounded, for the inversion recurrence p_k = p_(k-1) lam (1/k): one
 * multiplication instead of a division per step (same table on host and device) */
#if defined(__CUDACC__)
static __device__ const double beep_inv_dev[64] = {0.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 3.0, 1.0 / 4.0, 1.0 / 5.0, 1.0 / 6.0, 1.0 / 7.0, 1.0 / 8.0, 1.0 / 9.0, 1.0 / 10.0, 1.0 / 11.0, 1.0 / 12.0, 1.0 / 13.0, 1.0 / 14.0, 1.0 / 15.0, 1.0 / 16.0, 1.0 / 17.0, 1.0 / 18.0, 1.0 / 19.0, 1.0 / 20.0, 1.0 / 21.0, 1.0 / 22.0, 1.0 / 23.0, 1.0 / 24.0, 1.0 / 25.0, 1.0 / 26.0, 1.0 / 27.0, 1.0 / 28.0, 1.0 / 29.0, 1.0 / 30.0, 1.0 / 31.0, 1.0 / 32.0, 1.0 / 33.0, 1.0 / 34.0, 1.0 / 35.0, 1.0 / 36.0, 1.0 / 37.0, 1.0 / 38.0, 1.0 / 39.0, 1.0 / 40.0, 1.0 / 41.0, 1.0 / 42.0, 1.0 / 43.0, 1.0 / 44.0, 1.0 / 45.0, 1.0 / 46.0, 1.0 / 47.0, 1.0 / 48.0, 1.0 / 49.0, 1.0 / 50.0, 1.0 / 51.0, 1.0 / 52.0, 1.0 / 53.0, 1.0 / 54.0, 1.0 / 55.0, 1.0 / 56.0, 1.0 / 57.0, 1.0 / 58.0, 1.0 / 59.0, 1.0 / 60.0, 1.0 / 61.0, 1.0 / 62.0, 1.0 / 63.0};
#endif
static const double beep_inv_host[64] = {0.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 3.0, 1.0 / 4.0, 1.0 / 5.0, 1.0 / 6.0, 1.0 / 7.0, 1.0 / 8.0, 1.0 / 9.0, 1.0 / 10.0, 1.0 / 11.0, 1.0 / 12.0, 1.0 / 13.0, 1.0 / 14.0, 1.0 / 15.0, 1.0 / 16.0, 1.0 / 17.0, 1.0 / 18.0, 1.0 / 19.0, 1.0 / 20.0, 1.0 / 21.0, 1.0 / 22.0, 1.0 / 23.0, 1.0 / 24.0, 1.0 / 25.0, 1.0 / 26.0, 1.0 / 27.0, 1.0 / 28.0, 1.0 / 29.0, 1.0 / 30.0, 1.0 / 31.0, 1.0 / 32.0, 1.0 / 33.0, 1.0 / 34.0, 1.0 / 35.0, 1.0 / 36.0, 1.0 / 37.0, 1.0 / 38.0, 1.0 / 39.0, 1.0 / 40.0, 1.0 / 41.0, 1.0 / 42.0, 1.0 / 43.0, 1.0 / 44.0, 1.0 / 45.0, 1.0 / 46.0, 1.0 / 47.0, 1.0 / 48.0, 1.0 / 49.0, 1.0 / 50.0, 1.0 / 51.0, 1.0 / 52.0, 1.0 / 53.0, 1.0 / 54.0, 1.0 / 55.0, 1.0 / 56.0, 1.0 / 57.0, 1.0 / 58.0, 1.0 / 59.0, 1.0 / 60.0, 1.0 / 61.0, 1.0 / 62.0, 1.0 / 63.0};
#if defined(__CUDA_ARCH__)
#define BEEP_INV(k) beep_inv_dev[(k)]
#else
#define BEEP_INV(k) beep_inv_host[(k)]
#endif

/* PTRS squeeze failed: the exact acceptance test (Hormann 1993, step 3) */
BEEP_HD_COLD int beep_ptrs_accept(double lam, double a, double b, double us, double V, double kf)
{
#if defined(__CUDA_ARCH__)
	/* Device pre-test with the hardware-assisted log (libdevice, <= 1 ulp, ~1/3 of the instructions of
	 * the portable beep_log).  Both sides of the exact test below are sums of a few logarithms scaled
	 * by at most kf; the two evaluations can differ by a few ulps of the largest term (beep_log is
	 * itself within ~1 ulp of log, the Stirling tail is the same formula), i.e. by < 1e-14 * scale.
	 * Whenever the pre-test separates the two sides by more than 1e-9 * scale its verdict IS the
	 * exact test's verdict; otherwise (probability ~1e-6 per call) the exact test decides.  Results are
	 * therefore bit-identical to the host oracle, which always runs the exact test. */
	{
		const double inva = BR_ADD(1.1239, BR_DIV(1.1328, BR_SUB(b, 3.4)));
		const double tt = BR_ADD(BR_DIV(a, BR_MUL(us, us)), b);
		const double ll = log(lam);
		double lf;
		if (kf < 16.0) lf = beep_lfact(kf);                 /* table */
		else {
			const double ik = 1.0 / kf, ik2 = ik * ik;
			lf = (kf + 0.5) * log(kf) - kf + 9.189385332046727417803297364056e-01 +
			     ik * (8.333333333333333333e-02 + ik2 * (-2.777777777777777778e-03 + ik2 * (7.936507936507936508e-04 + ik2 * -5.952380952380952381e-04)));
		}
		const double lhs = log(V * inva / tt);                /* = log V + log(1/alpha) - log(a/us^2 + b): one logarithm, not three */
		const double klam = kf * ll;
		const double rhs = klam - lam - lf;
		const double scale = fabs(klam) + lam + fabs(lf) + fabs(lhs) + 1.0;
		const double diff = lhs - rhs;
		if (fabs(diff) > 1e-9 * scale) return diff <= 0.0;
	}
#endif
	double invalpha = BR_ADD(1.1239, BR_DIV(1.1328, BR_SUB(b, 3.4)));
	double loglam = beep_log(lam);
	double lhs_c = beep_log(invalpha);
	double lhs = BR_SUB(BR_ADD(beep_log(V), lhs_c), beep_log(BR_ADD(BR_DIV(a, BR_MUL(us, us)), b)));
	double rhs = BR_SUB(BR_SUB(BR_MUL(kf, loglam), lam), beep_lfact(kf));
	return lhs <= rhs;
}

/* PTRS transformed rejection (Hormann 1993) for lam >= 10 */
BEEP_HD int beep_poisson_ptrs(BeepStream *s, double lam)
{
	double slam = BR_SQRT(lam);
	double b = BR_ADD(0.931, BR_MUL(2.53, slam));
	double a = BR_ADD(-0.059, BR_MUL(0.02483, b));
	double vr = BR_SUB(0.9277, BR_DIV(3.6224, BR_SUB(b, 2.0)));
	for (int it = 0; it < 100000; it++) {
		double U = BR_SUB(beep_next_uniform(s), 0.5);
		double V = beep_next_uniform(s);
		double us = BR_SUB(0.5, BR_FABS(U));
		double kf = BR_FLOOR(BR_ADD(BR_ADD(BR_MUL(BR_ADD(BR_DIV(BR_MUL(2.0, a), us), b), U), lam), 0.43));
		if (us >= 0.07 && V <= vr) return (int)kf;
		if (kf < 0.0 || (us < 0.013 && V > us)) continue;
		if (beep_ptrs_accept(lam, a, b, us, V, kf)) return (int)kf;
	}
	return (int)lam;
}
#if defined(__CUDACC__)
/* Device: the FIRST PTRS attempt of a fresh stream, squeeze only.  Returns 1 and the draw when the
 * squeeze accepts (the same value beep_poisson_ptrs returns from its first iteration); 0 when the
 * attempt needs the exact test or is rejected -- the caller then runs beep_poisson_ptrs_fresh,
 * which repeats this attempt and carries on.  Lets a warp do the cheap, common part of all its
 * cells first and the expensive logarithms of the rest together. */
static __device__ __forceinline__ int beep_poisson_ptrs_try(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double lam, int *out, const uint32_t *rk = 0)
{
	BeepStream st;
	st.key[0] = k0; st.key[1] = k1; st.ctr[0] = c0; st.ctr[1] = c1; st.ctr[2] = c2; st.ctr[3] = c3; st.have = 0; st.rk = rk;
	double slam = BR_SQRT(lam);
	double b = BR_ADD(0.931, BR_MUL(2.53, slam));
	double a = BR_ADD(-0.059, BR_MUL(0.02483, b));
	double vr = BR_SUB(0.9277, BR_DIV(3.6224, BR_SUB(b, 2.0)));
	double U = BR_SUB(beep_next_uniform(&st), 0.5);
	double V = beep_next_uniform(&st);
	double us = BR_SUB(0.5, BR_FABS(U));
	double kf = BR_FLOOR(BR_ADD(BR_ADD(BR_MUL(BR_ADD(BR_DIV(BR_MUL(2.0, a), us), b), U), lam), 0.43));
	if (us >= 0.07 && V <= vr) { *out = (int)kf; return 1; }
	return 0;
}
/* Device: a fresh stream (nothing read yet) runs PTRS out of line, so that kernels dominated
 * by small means keep a small hot loop; the caller's copy of the stream is not advanced --
 * on the device a stream serves exactly one draw. */
static __device__ __noinline__ int beep_poisson_ptrs_fresh(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double lam)
{
	BeepStream st;
	st.key[0] = k0; st.key[1] = k1; st.ctr[0] = c0; st.ctr[1] = c1; st.ctr[2] = c2; st.ctr[3] = c3; st.have = 0; st.rk = 0;
	return beep_poisson_ptrs(&st, lam);
}
#endif

/* sequential-search inversion for 0 < lam < 10 from the uniform u: the exact search (correctly rounded FP64
 * operations in a fixed order -- what host and device must agree on) */
BEEP_HD int beep_poisson_inv_search(double u, double lam)
{
	double p = beep_exp(BR_SUB(0.0, lam));
	double cdf = p;
	int k = 0;
	double ik = BEEP_INV(1);                       /* 1/(k+1), fetched one step ahead of its use */
	while (u > cdf && k < 1000) {
		k++;
		double ikn = BEEP_INV(k < 62 ? k + 1 : 63);
		p = (k < 64) ? BR_MUL(BR_MUL(p, lam), ik) : BR_DIV(BR_MUL(p, lam), (double)k);
		cdf = BR_ADD(cdf, p);
		ik = ikn;
	}
	return k;
}
#if defined(__CUDA_ARCH__) && !defined(BEEP_NO_F32_PRESEARCH)
static __device__ __noinline__ int beep_poisson_inv_search_cold(double u, double lam) { return beep_poisson_inv_search(u, lam); }
#endif

BEEP_HD int beep_poisson_inv_u(double u, double lam)
{
	/* exp(-lam) >= 1 - lam, and beep_exp is good to ~1 ulp: a uniform below
	 * (1 - lam) - 4.5e-16 yields k = 0 whatever beep_exp rounds to, so the
	 * shortcut never changes a draw (it only skips the exp) */
	if (u <= BR_SUB(BR_SUB(1.0, lam), 4.5e-16)) return 0;
#if defined(__CUDA_ARCH__) && !defined(BEEP_NO_F32_PRESEARCH)
	/* Device pre-search in single precision (hardware exp2 and reciprocal): the exact search returns the first k with
	 * u <= cdf_k, its cdf_k within ~1e-14 of the true Poisson sums.  The single-precision sums c_k below carry a relative
	 * error of at most 1.6e-6 (__expf: 2 + |1.173 x| ulp at |x| < 10) + 6e-7 (lam rounded to float) + 4.3e-7 k (three
	 * roundings and an approximate reciprocal per step, the summation) < 2.2e-6 + 5e-7 k, and u rounds by 6e-8.  When u is
	 * further than tol = 5e-6 + 1e-6 k (more than twice that bound; cdf <= 1) from BOTH c_(k-1) and c_k, the exact search
	 * sees cdf_(k-1) < u <= cdf_k as well (its sums are monotone, so the earlier comparisons follow) and returns this k:
	 * the draw is unchanged.  Otherwise (probability ~ 4 tol per draw) the exact search decides, out of line. */
	{
		const float uf = __double2float_rn(u), lf = __double2float_rn(lam);
		float p = __expf(-lf), cdf = p, prev = -1.0f, kf = 0.0f;
		while (uf > cdf && kf < 60.0f) {
			kf += 1.0f;
			float rk;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rk) : "f"(kf));      /* <= 1 ulp */
			p = __fmul_rn(__fmul_rn(p, lf), rk);
			prev = cdf;
			cdf = __fadd_rn(cdf, p);
		}
		const float tol = __fadd_rn(5e-6f, __fmul_rn(1e-6f, kf));
		if (__fsub_rn(cdf, uf) > tol && __fsub_rn(uf, prev) > tol) return (int)kf;
		return beep_poisson_inv_search_cold(u, lam);
	}
#else
	return beep_poisson_inv_search(u, lam);
#endif
}

/* the first two uniforms of a fresh per-cell stream (one Philox block): what the first two calls of
 * beep_next_uniform return after beep_stream_init_cell */
BEEP_HD void beep_cell_uniforms(uint64_t seed, uint64_t stream_id, uint64_t cell, double *u1, double *u2)
{
	uint32_t key[2], ctr[4], buf[4];
	key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
	ctr[0] = (uint32_t)(cell >> 32) << 16; ctr[1] = (uint32_t)cell;
	ctr[2] = (uint32_t)stream_id; ctr[3] = (uint32_t)(stream_id >> 32);
	beep_philox4x32_10(ctr, key, buf);
	uint64_t x = (((uint64_t)buf[0] << 32) | (uint64_t)buf[1]) >> 12;
	*u1 = BR_MUL(BR_ADD((double)x, 0.5), 2.220446049250313080847263336181640625e-16);
	x = (((uint64_t)buf[2] << 32) | (uint64_t)buf[3]) >> 12;
	*u2 = BR_MUL(BR_ADD((double)x, 0.5), 2.220446049250313080847263336181640625e-16);
}

/* the FIRST PTRS attempt from the stream's first two uniforms, squeeze only: returns 1 and the draw when the
 * squeeze accepts (the value beep_poisson_ptrs returns from its first iteration), 0 when the attempt needs
 * the exact test or is rejected */
BEEP_HD int beep_poisson_ptrs_try_uv(double u1, double u2, double lam, int *out)
{
	double slam = BR_SQRT(lam);
	double b = BR_ADD(0.931, BR_MUL(2.53, slam));
	double a = BR_ADD(-0.059, BR_MUL(0.02483, b));
	double vr = BR_SUB(0.9277, BR_DIV(3.6224, BR_SUB(b, 2.0)));
	double U = BR_SUB(u1, 0.5);
	double V = u2;
	double us = BR_SUB(0.5, BR_FABS(U));
	double kf = BR_FLOOR(BR_ADD(BR_ADD(BR_MUL(BR_ADD(BR_DIV(BR_MUL(2.0, a), us), b), U), lam), 0.43));
	if (us >= 0.07 && V <= vr) { *out = (int)kf; return 1; }
	return 0;
}

BEEP_HD int beep_poisson(BeepStream *s, double lam)
{
	if (!(lam > 0.0)) return 0;
	if (lam < BEEP_POISSON_INV_MAX) return beep_poisson_inv_u(beep_next_uniform(s), lam);
#if defined(__CUDA_ARCH__)
	if (s->have == 0) return beep_poisson_ptrs_fresh(s->key[0], s->key[1], s->ctr[0], s->ctr[1], s->ctr[2], s->ctr[3], lam);
#endif
	return beep_poisson_ptrs(s, lam);
}

/* ------------------------------------------------------------------------ */
/* Binomial                                                                  */
/* ------------------------------------------------------------------------ */
/* BTRS squeeze failed: the exact acceptance test (Hormann 1993) */
BEEP_HD_COLD int beep_btrs_accept(double dn, double q, double omq, double spq, double a, double b, double us, double V, double kf)
{
	double alpha = BR_MUL(BR_ADD(2.83, BR_DIV(5.1, b)), spq);
	double lpq = beep_log(BR_DIV(q, omq));
	double m = BR_FLOOR(BR_MUL(BR_ADD(dn, 1.0), q));
	double h = BR_ADD(beep_lfact(m), beep_lfact(BR_SUB(dn, m)));
	double lv = beep_log(BR_DIV(BR_MUL(V, alpha), BR_ADD(BR_DIV(a, BR_MUL(us, us)), b)));
	double rhs = BR_ADD(BR_SUB(BR_SUB(h, beep_lfact(kf)), beep_lfact(BR_SUB(dn, kf))), BR_MUL(BR_SUB(kf, m), lpq));
	return lv <= rhs;
}

/* BINV start value (1-q)^n = exp(n log(1-q)) */
BEEP_HD_COLD double beep_binv_r0(double dn, double omq) { return beep_exp(BR_MUL(dn, beep_log(omq))); }

BEEP_HD int beep_binomial_body(BeepStream *s, double p, int n)
{
	if (n <= 0 || !(p > 0.0)) return 0;
	if (p >= 1.0) return n;
	int flip = p > 0.5;
	double q = flip ? BR_SUB(1.0, p) : p;
	double dn = (double)n;
	double omq = BR_SUB(1.0, q);
	int k;
	if (BR_MUL(dn, q) < 10.0) {
		/* BINV */
		double u = beep_next_uniform(s);
		k = 0;
		/* (1-q)^n >= 1 - n q; the computed r = exp(n log(1-q)) is within ~n q ulps of it
		 * (n q < 10): below (1 - n q) - 1e-14 the answer is 0 whatever r rounds to */
		if (!(u <= BR_SUB(BR_SUB(1.0, BR_MUL(dn, q)), 1e-14))) {
			double sr = BR_DIV(q, omq);
			double a = BR_MUL(BR_ADD(dn, 1.0), sr);
			double r = beep_binv_r0(dn, omq);
			while (u > r && k < n) {
				u = BR_SUB(u, r);
				k++;
				r = BR_MUL(r, BR_SUB(BR_DIV(a, (double)k), sr));
				if (!(r > 0.0)) break;
			}
		}
	} else {
		/* BTRS */
		double spq = BR_SQRT(BR_MUL(BR_MUL(dn, q), omq));
		double b = BR_ADD(1.15, BR_MUL(2.53, spq));
		double a = BR_ADD(BR_ADD(-0.0873, BR_MUL(0.0248, b)), BR_MUL(0.01, q));
		double c = BR_ADD(BR_MUL(dn, q), 0.5);
		double vr = BR_SUB(0.92, BR_DIV(4.2, b));
		k = (int)BR_MUL(dn, q);
		for (int it = 0; it < 100000; it++) {
			double U = BR_SUB(beep_next_uniform(s), 0.5);
			double V = beep_next_uniform(s);
			double us = BR_SUB(0.5, BR_FABS(U));
			double kf = BR_FLOOR(BR_ADD(BR_MUL(BR_ADD(BR_DIV(BR_MUL(2.0, a), us), b), U), c));
			if (kf < 0.0 || kf > dn) continue;
			if (us >= 0.07 && V <= vr) { k = (int)kf; break; }
			if (beep_btrs_accept(dn, q, omq, spq, a, b, us, V, kf)) { k = (int)kf; break; }
		}
	}
	return flip ? n - k : k;
}

#if defined(__CUDACC__)
/* Device: the whole binomial draw of a fresh stream runs out of line (see beep_poisson_ptrs_fresh) */
static __device__ __noinline__ int beep_binomial_fresh(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double p, int n)
{
	BeepStream st;
	st.key[0] = k0; st.key[1] = k1; st.ctr[0] = c0; st.ctr[1] = c1; st.ctr[2] = c2; st.ctr[3] = c3; st.have = 0; st.rk = 0;
	return beep_binomial_body(&st, p, n);
}
#endif

BEEP_HD int beep_binomial(BeepStream *s, double p, int n)
{
#if defined(__CUDA_ARCH__)
	if (s->have == 0) return beep_binomial_fresh(s->key[0], s->key[1], s->ctr[0], s->ctr[1], s->ctr[2], s->ctr[3], p, n);
#endif
	return beep_binomial_body(s, p, n);
}

#endif /* BEEP_RNG_H */
