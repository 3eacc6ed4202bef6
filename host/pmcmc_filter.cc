// pmcmc_filter.cc -- C++ host example above the C ABI: the bootstrap particle filter that
// PMCMC::get_al evaluates for every proposal (pmcmc.cc:180-268), driven from a model file written
// by the reference-side adapter.  usage: pmcmc_filter model.bpk nparticle [nrep] [seed]
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "beep_b200.hpp"
#include "bpk.hpp"

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s model.bpk nparticle [nrep] [seed]\n", argv[0]); return 2; }
	try {
		beep::ModelFile mf(argv[1]);
		const int P = atoi(argv[2]), nrep = argc > 3 ? atoi(argv[3]) : 3;
		const uint64_t seed = argc > 4 ? strtoull(argv[4], nullptr, 10) : 1;
		auto th = mf.arrays.find("theta");
		if (th == mf.arrays.end()) throw std::runtime_error("model file has no 'theta' array");
		const double *theta0 = (const double *)th->second.data.data();
		const int np = mf.desc.nparam;
		std::vector<double> theta((size_t)P * np);
		for (int p = 0; p < P; p++) for (int k = 0; k < np; k++) theta[(size_t)p * np + k] = theta0[k];   // all particles share theta (pmcmc.cc:186)
		beep::Engine eng(mf.desc, 0);
		beep::BatchState particles(eng, P);
		for (int r = 0; r < nrep; r++) {
			auto t0 = std::chrono::steady_clock::now();
			double ll = beep::particle_filter(eng, particles, theta, seed + r);
			double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			printf("filter %d: P=%d sections=%d log-likelihood=%.10f  %.2f ms\n", r, P, eng.nsection(), ll, ms);
		}
	} catch (const std::exception &e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
