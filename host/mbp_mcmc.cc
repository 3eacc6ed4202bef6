// mbp_mcmc.cc -- C++ host example above the C ABI: MCMC with model-based proposals, the inner loop of the
// reference's mcmcmbp / mc3 / abcmbp / pas drivers (Mbp::mbp_proposal, mbp.cc:426-456), for a batch of
// independent chains on one GPU.  Per sweep and free parameter: a random-walk proposal of that parameter
// for every chain (param_prop.cc single-parameter updates), State::set_param + Mbp::mbp for the whole batch
// (beep_set_params, beep_mbp), the observation likelihood of the proposed states (beep_obs_loglik), the
// acceptance probability of Mbp::get_al in INVT mode (mbp.cc:334-361: al = exp(-invT (EF_p - EF_i)/2 +
// Pr_p - Pr_i), EF = -2 L, uniform priors) and Mbp::swap_initial_propose_state for the accepted chains
// (beep_accept); then one Mbp::mbp_slicetime-style proposal (parameters fixed, a random fraction of the
// (step, area) blocks re-simulated) that refreshes the event sequence.  Step sizes adapt during burn-in (utils.cc:619-637 style) towards 33 % acceptance.
//
// usage: mbp_mcmc model.bpk nchain nburn nsamp seed  k:lo:hi[:sigma] ...     (k = parameter index, uniform prior on [lo,hi])
// prints: "param k mean M sd S" for every free parameter, pooled over chains and post-burn-in sweeps.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "beep_b200.hpp"
#include "bpk.hpp"

struct Free { int k; double lo, hi, sigma; long ntr = 0, nac = 0; };

int main(int argc, char **argv)
{
	if (argc < 7) { fprintf(stderr, "usage: %s model.bpk nchain nburn nsamp seed k:lo:hi[:sigma] ...\n", argv[0]); return 2; }
	try {
		beep::ModelFile mf(argv[1]);
		const int P = atoi(argv[2]), nburn = atoi(argv[3]), nsamp = atoi(argv[4]);
		const uint64_t seed = strtoull(argv[5], nullptr, 10);
		std::vector<Free> fr;
		for (int i = 6; i < argc; i++) {
			Free f; f.sigma = 0;
			if (sscanf(argv[i], "%d:%lf:%lf:%lf", &f.k, &f.lo, &f.hi, &f.sigma) < 3) throw std::runtime_error("bad free-parameter spec");
			if (f.sigma <= 0) f.sigma = 0.05 * (f.hi - f.lo);
			fr.push_back(f);
		}
		auto th = mf.arrays.find("theta");
		if (th == mf.arrays.end()) throw std::runtime_error("model file has no 'theta' array");
		const double *theta0 = (const double *)th->second.data.data();
		const int np = mf.desc.nparam, T = mf.desc.T;
		const double invT = 1.0;

		std::mt19937_64 gen(seed);
		std::normal_distribution<double> nd(0, 1);
		std::uniform_real_distribution<double> ud(0, 1);
		std::vector<double> theta((size_t)P * np), prop_theta((size_t)P * np);
		for (int c = 0; c < P; c++) {
			for (int k = 0; k < np; k++) theta[(size_t)c * np + k] = theta0[k];
			for (auto &f : fr) theta[(size_t)c * np + f.k] = f.lo + (0.25 + 0.5 * ud(gen)) * (f.hi - f.lo);   // over-dispersed starts
		}
		std::vector<uint64_t> ids(P);
		for (int c = 0; c < P; c++) ids[c] = (uint64_t)c;

		beep::Engine eng(mf.desc, 0);
		const unsigned flags = BEEP_KEEP_TRANSNUM | BEEP_KEEP_TRANSMEAN;
		beep::BatchState initial(eng, P, flags), propose(eng, P, flags);
		initial.set_param(theta);
		initial.simulate(0, T, seed, ids.data());
		std::vector<double> Li = initial.obs_loglik();
		propose.set_param(theta);
		propose.simulate(0, T, seed + 1, ids.data());          // installs the stream ids of the proposal batch

		std::vector<double> sum(fr.size(), 0), sum2(fr.size(), 0);
		long nacc = 0;
		std::vector<uint8_t> accept(P), mask((size_t)P * T * mf.desc.A);
		double sim_frac = 0.05;
		long st_ntr = 0, st_nac = 0;
		auto t0 = std::chrono::steady_clock::now();
		uint64_t it_seed = seed * 1000003ull + 17;
		for (int sweep = 0; sweep < nburn + nsamp; sweep++) {
			for (size_t j = 0; j < fr.size(); j++) {
				Free &f = fr[j];
				prop_theta = theta;
				std::vector<uint8_t> inb(P, 1);
				for (int c = 0; c < P; c++) {
					double v = theta[(size_t)c * np + f.k] + f.sigma * nd(gen);
					if (v < f.lo || v > f.hi) inb[c] = 0;                      // outside the prior: rejected, the chain keeps its state
					else prop_theta[(size_t)c * np + f.k] = v;
				}
				propose.set_param(prop_theta);
				propose.mbp(initial, ++it_seed);
				std::vector<double> Lp = propose.obs_loglik();
				for (int c = 0; c < P; c++) {
					double al = inb[c] ? std::exp(invT * (Lp[c] - Li[c])) : 0.0;   // exp(-invT (EF_p - EF_i) / 2)
					accept[c] = (ud(gen) < al) ? 1 : 0;
					f.ntr++;
					if (accept[c]) { f.nac++; theta[(size_t)c * np + f.k] = prop_theta[(size_t)c * np + f.k]; Li[c] = Lp[c]; }
				}
				initial.accept_from(propose, accept);
				if (sweep < nburn && (sweep + 1) % 10 == 0) {               // adapt the jump size during burn-in
					double ar = (double)f.nac / (double)f.ntr;
					f.sigma *= (ar > 0.33) ? 1.15 : 0.87;
					if (f.sigma > f.hi - f.lo) f.sigma = f.hi - f.lo;
					f.ntr = f.nac = 0;
				}
			}
			{
				/* Mbp::mbp_slicetime (mbp.cc:687-713): parameters fixed, a random fraction of the (step, area)
				 * blocks is re-simulated (SIMU), the rest follows the model-based proposal; refreshes the
				 * event sequence independently of the parameters.  The fraction adapts during burn-in. */
				const int A = mf.desc.A;
				for (size_t i = 0; i < mask.size(); i++) mask[i] = (ud(gen) < sim_frac) ? 1 : 0;
				propose.set_param(theta);
				propose.mbp(initial, ++it_seed, mask.data());
				std::vector<double> Lp = propose.obs_loglik();
				for (int c = 0; c < P; c++) {
					double al = std::exp(invT * (Lp[c] - Li[c]));
					accept[c] = (ud(gen) < al) ? 1 : 0;
					st_ntr++;
					if (accept[c]) { st_nac++; Li[c] = Lp[c]; }
				}
				initial.accept_from(propose, accept);
				if (sweep < nburn && (sweep + 1) % 10 == 0) {
					sim_frac *= ((double)st_nac / st_ntr > 0.33) ? 1.15 : 0.87;
					if (sim_frac > 1) sim_frac = 1;
					st_ntr = st_nac = 0;
				}
				(void)A;
			}
			if (sweep >= nburn) {
				for (size_t j = 0; j < fr.size(); j++)
					for (int c = 0; c < P; c++) { double v = theta[(size_t)c * np + fr[j].k]; sum[j] += v; sum2[j] += v * v; }
				nacc += P;
			}
		}
		double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		for (size_t j = 0; j < fr.size(); j++) {
			double m = sum[j] / nacc, v = sum2[j] / nacc - m * m;
			printf("param %d mean %.6f sd %.6f acceptance %.3f sigma %.4f\n", fr[j].k, m, std::sqrt(v > 0 ? v : 0), fr[j].ntr ? (double)fr[j].nac / fr[j].ntr : 0.0, fr[j].sigma);
		}
		printf("slice-time sim_frac %.4f acceptance %.3f\n", sim_frac, st_ntr ? (double)st_nac / st_ntr : 0.0);
		printf("chains %d sweeps %d proposals %ld seconds %.3f proposals_per_s %.0f\n", P, nburn + nsamp, (long)P * (nburn + nsamp) * (long)fr.size(), sec,
		       (double)P * (nburn + nsamp) * fr.size() / sec);
	} catch (const std::exception &e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
