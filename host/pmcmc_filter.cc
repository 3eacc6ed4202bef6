// pmcmc_filter.cc -- C++ host example above the C ABI: the bootstrap particle filter that
// PMCMC::get_al evaluates for every proposal (pmcmc.cc:180-268), driven from a model file written
// by the reference-side adapter.  usage: pmcmc_filter model.bpk nparticle_total [nrep] [seed]
//
// One process per GPU: with WORLD_SIZE > 1 (RANK / WORLD_SIZE / LOCAL_RANK as torchrun or mpirun export them) the
// particles are sharded over the ranks as the reference shards them over MPI cores (mpi.cc:98-101), the weights
// are all-gathered and the ancestors moved by the library's NCCL collectives (beep::Comm); the 128 bytes of the
// ncclUniqueId travel through a file (BEEP_NCCL_ID_FILE, default /tmp/beep_filter_id_<MASTER_PORT>).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>

#include "beep_b200.hpp"
#include "bpk.hpp"

static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

static std::vector<unsigned char> exchange_id(int rank)
{
	const char *port = getenv("MASTER_PORT");
	std::string file = getenv("BEEP_NCCL_ID_FILE") ? getenv("BEEP_NCCL_ID_FILE") : std::string("/tmp/beep_filter_id_") + (port ? port : "0");
	std::vector<unsigned char> id(128);
	if (rank == 0) {
		id = beep::Comm::unique_id();
		FILE *f = fopen((file + ".tmp").c_str(), "wb");
		if (!f) throw std::runtime_error("cannot write " + file);
		fwrite(id.data(), 1, 128, f); fclose(f);
		rename((file + ".tmp").c_str(), file.c_str());
	} else {
		for (int spin = 0;; spin++) {
			FILE *f = fopen(file.c_str(), "rb");
			if (f) { size_t n = fread(id.data(), 1, 128, f); fclose(f); if (n == 128) break; }
			if (spin > 60000) throw std::runtime_error("timed out waiting for " + file);
			std::this_thread::sleep_for(std::chrono::milliseconds(10));
		}
	}
	return id;
}

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s model.bpk nparticle_total [nrep] [seed]\n", argv[0]); return 2; }
	try {
		const int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1), dev = env_int("LOCAL_RANK", 0);
		beep::ModelFile mf(argv[1]);
		const int Ptot = atoi(argv[2]), nrep = argc > 3 ? atoi(argv[3]) : 3;
		if (Ptot % nranks) throw std::runtime_error("the number of particles must be a multiple of the number of ranks");
		const int P = Ptot / nranks;
		const uint64_t seed = argc > 4 ? strtoull(argv[4], nullptr, 10) : 1;
		auto th = mf.arrays.find("theta");
		if (th == mf.arrays.end()) throw std::runtime_error("model file has no 'theta' array");
		const double *theta0 = (const double *)th->second.data.data();
		const int np = mf.desc.nparam;
		std::vector<double> theta((size_t)P * np);
		for (int p = 0; p < P; p++) for (int k = 0; k < np; k++) theta[(size_t)p * np + k] = theta0[k];   // all particles share theta (pmcmc.cc:186)
		beep::Engine eng(mf.desc, dev);
		std::unique_ptr<beep::Comm> comm;
		if (nranks > 1) {
			comm.reset(new beep::Comm(eng, exchange_id(rank), nranks, rank));
			comm->allgather({0.0});                                          // everyone is here: the file can go
			if (rank == 0) remove((getenv("BEEP_NCCL_ID_FILE") ? std::string(getenv("BEEP_NCCL_ID_FILE")) : std::string("/tmp/beep_filter_id_") + (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0")).c_str());
		}
		beep::BatchState particles(eng, P);
		for (int r = 0; r < nrep; r++) {
			auto t0 = std::chrono::steady_clock::now();
			double ll = beep::particle_filter(eng, particles, theta, seed + r, 1.0, comm.get());
			double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			if (rank == 0) printf("filter %d: P=%d ranks=%d sections=%d log-likelihood=%.10f  %.2f ms\n", r, Ptot, nranks, eng.nsection(), ll, ms);
		}
	} catch (const std::exception &e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
