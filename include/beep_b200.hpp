// beep_b200.hpp -- header-only C++ wrapper over the C ABI (include/beep_b200.h).
//
// The reference's drivers hold `State` / `vector<State> particle` / `vector<Particle> part`
// by value and call their members directly (pmcmc.cc:60-66,180-207; abcsmc.cc:42-71;
// mbp.cc:99-133).  These RAII classes are the batched counterparts a driver holds instead:
//   beep::Engine      <- Details + Data + Model tables on one GPU   (main.cc:125-151)
//   beep::BatchState  <- vector<State>                              (state.hh:11-133)
//   beep::Comm        <- Mpi, for the particle plumbing of the hot path (mpi.hh:16-60): one per GPU / rank
// Errors become exceptions carrying beep_last_error() (the reference exits, utils.cc:296-316).
#ifndef BEEP_B200_HPP
#define BEEP_B200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "beep_b200.h"

namespace beep {

struct Error : std::runtime_error {
	int code;
	Error(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};
inline void check(int rc) { if (rc != BEEP_OK) throw Error(rc, beep_last_error()); }

class Engine {
public:
	Engine(const BeepModelDesc &desc, int device = 0)
	{
		if (beep_desc_size() != (int)sizeof(BeepModelDesc)) throw Error(BEEP_ERR_INVALID, "BeepModelDesc of this program and of libbeep_b200.so differ (rebuild against the library's beep_model_fields.def)");
		check(beep_engine_create(&desc, device, &h_));
	}
	~Engine() { beep_engine_destroy(h_); }
	Engine(const Engine &) = delete;
	Engine &operator=(const Engine &) = delete;
	BeepEngine *handle() const { return h_; }
	int nsection() const { return beep_num_sections(h_); }
	int ntable() const { int v = 0; check(beep_engine_dims(h_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &v)); return v; }
	/// A*J*D: entries of one time step of Particle::transnum
	size_t cells_per_step() const { int A = 0, J = 0, D = 0; check(beep_engine_dims(h_, nullptr, &A, nullptr, &J, &D, nullptr, nullptr, nullptr)); return (size_t)A * J * D; }
	void section_range(int s, int &ti, int &tf) const { check(beep_section_range(h_, s, &ti, &tf)); }
	/// ABCSMC::calculate_particle_weight's sum over the previous generation (abcsmc.cc:264-268) for all proposals:
	/// sums[i] = SUM_j w_last[j] exp(MVN::probability(prop_i, last_j)); parameters are rows of nparam doubles
	std::vector<double> mvn_kernel_sums(const std::vector<int> &var, const std::vector<double> &inv_M, double size, int nparam,
	                                    const std::vector<double> &param_prop, const std::vector<double> &param_last, const std::vector<double> &w_last) const
	{
		const int N = (int)(param_prop.size() / nparam), Nl = (int)w_last.size();
		std::vector<double> sums(N);
		check(beep_mvn_kernel_sums(h_, (int)var.size(), nparam, var.data(), inv_M.data(), size, param_prop.data(), N, param_last.data(), w_last.data(), Nl, sums.data()));
		return sums;
	}
private:
	BeepEngine *h_ = nullptr;
};

/// P particles in HBM.  Method names follow State / ObservationModel / Mbp of the reference.
class BatchState {
public:
	BatchState(Engine &eng, int nparticle, unsigned flags = 0) : eng_(eng) { check(beep_batch_create(eng.handle(), nparticle, flags, &h_)); }
	~BatchState() { beep_batch_destroy(h_); }
	BatchState(const BatchState &) = delete;
	BatchState &operator=(const BatchState &) = delete;
	BeepBatch *handle() const { return h_; }
	int size() const { return beep_batch_size(h_); }

	/// State::set_param (state.cc:66-91) for every particle; theta is [P][nparam]
	void set_param(const std::vector<double> &theta) { check(beep_set_params(h_, theta.data())); }
	/// State::simulate(ti,tf) (state.cc:596-638)
	void simulate(int ti, int tf, uint64_t seed, const uint64_t *stream_ids = nullptr) { check(beep_simulate(h_, ti, tf, seed, stream_ids)); }
	/// ObservationModel::calculate / calculate_section (obsmodel.cc:22-58); EF = -2 L (state.cc:95-98)
	std::vector<double> obs_loglik(int section = -1)
	{
		std::vector<double> L(size());
		check(beep_obs_loglik(h_, section, L.data(), nullptr));
		return L;
	}
	/// State::likelihood (state.cc:812-826)
	std::vector<double> likelihood()
	{
		std::vector<double> Li(size());
		check(beep_latent_loglik(h_, Li.data()));
		return Li;
	}
	/// State::likelihood_approx (state_mvn_approx.cc:11-91) of every particle at inverse temperature invT
	std::vector<double> likelihood_approx(double invT)
	{
		std::vector<double> L(size());
		check(beep_likelihood_approx(h_, invT, L.data()));
		return L;
	}
	/// Mbp::mbp (mbp.cc:252-330): this batch is the proposal, coupled slot by slot to `initial`
	void mbp(BatchState &initial, uint64_t seed, const uint8_t *simu_mask = nullptr) { check(beep_mbp(initial.h_, h_, simu_mask, BEEP_INF_DIF_UPDATE, seed, nullptr)); }
	/// Mbp::swap_initial_propose_state (mbp.cc:243-248) for the accepted slots
	void accept_from(BatchState &proposal, const std::vector<uint8_t> &accept) { check(beep_accept(h_, proposal.h_, accept.data())); }
	/// State::initialise_from_particle (state.cc:294-322)
	void replay() { check(beep_replay(h_)); }
	/// Mpi::copy_particles / end_sec_swap within one GPU (mpi.cc:81-165,215-327)
	void permute(const std::vector<uint32_t> &src_of, int what) { check(beep_permute(h_, src_of.data(), what)); }
private:
	Engine &eng_;
	BeepBatch *h_ = nullptr;
};

/// ABCMBP::EF_cutoff (abcmbp.cc:135-217): cut-off and copy list from the all-gathered error functions
struct CutoffDecision { double EFcut, EFmin, EFmax; std::vector<uint32_t> partcopy; };
inline CutoffDecision ef_cutoff(const std::vector<double> &EF_all)
{
	CutoffDecision d;
	d.partcopy.resize(EF_all.size());
	check(beep_ef_cutoff(EF_all.data(), (int)EF_all.size(), &d.EFcut, &d.EFmin, &d.EFmax, d.partcopy.data()));
	return d;
}
/// PAS::bootstrap (pas.cc:127-209): inverse-temperature step and copy list; invT_final < 0 = unset
struct PasDecision { double DinvT, EFmin, EFmax; std::vector<uint32_t> partcopy; };
inline PasDecision pas_bootstrap(const std::vector<double> &EF_all, int nrun, int npart, double invT, double quench_factor, double invT_final, uint64_t seed, int gen)
{
	PasDecision d;
	d.partcopy.resize(EF_all.size());
	check(beep_pas_bootstrap(EF_all.data(), nrun, npart, invT, quench_factor, invT_final, seed, gen, &d.DinvT, &d.EFmin, &d.EFmax, d.partcopy.data()));
	return d;
}
/// MC3::swap_states (mc3.cc:153-217): new chain order and the proposal / acceptance counts to add per position
struct SwapDecision { std::vector<uint32_t> order, ntr, nac; };
inline SwapDecision mc3_swap(const std::vector<uint32_t> &run_all, const std::vector<double> &invT_all, const std::vector<double> &EF_all, uint64_t seed, int sweep)
{
	SwapDecision d;
	const int n = (int)EF_all.size();
	d.order.resize(n); d.ntr.resize(n); d.nac.resize(n);
	check(beep_mc3_swap(run_all.data(), invT_all.data(), EF_all.data(), n, seed, sweep, d.order.data(), d.ntr.data(), d.nac.data()));
	return d;
}

/// One communicator per GPU / rank (NCCL inside the library, particles move HBM to HBM).  Global particle i lives
/// on rank i / N, the reference's core = i/N (mpi.cc:98-101).  `id128` are the 128 bytes of an ncclUniqueId written
/// by Comm::unique_id() on one rank and handed to the others by whatever the host has (MPI_Bcast, a file, ...).
class Comm {
public:
	static std::vector<unsigned char> unique_id() { std::vector<unsigned char> id(128); check(beep_comm_unique_id(id.data())); return id; }
	Comm(Engine &eng, const std::vector<unsigned char> &id128, int nranks, int rank) { check(beep_comm_create(eng.handle(), id128.data(), nranks, rank, &h_)); }
	~Comm() { beep_comm_destroy(h_); }
	Comm(const Comm &) = delete;
	Comm &operator=(const Comm &) = delete;
	BeepComm *handle() const { return h_; }
	int rank() const { return beep_comm_rank(h_); }
	int size() const { return beep_comm_size(h_); }
	/// mpi.gather + bcast of per-particle scalars (abcmbp.cc:141, pas.cc:133, pmcmc.cc:219,263, mc3.cc:170-175)
	std::vector<double> allgather(const std::vector<double> &local)
	{
		std::vector<double> all(local.size() * (size_t)size());
		check(beep_allgather_scalar(h_, local.data(), (int)local.size(), all.data()));
		return all;
	}
	/// Mpi::copy_particles (mpi.cc:81-165): global particle i continues as a copy of partcopy_all[i] (0xFFFFFFFF = keep)
	void copy_particles(BatchState &batch, const std::vector<uint32_t> &partcopy_all) { check(beep_copy_particles(h_, batch.handle(), partcopy_all.data(), (int)partcopy_all.size())); }
	/// Mpi::end_sec_swap (mpi.cc:215-327): particle p continues from the section state of backpart_all[p]
	void end_sec_swap(BatchState &batch, const std::vector<uint32_t> &backpart_all) { check(beep_end_sec_swap(h_, batch.handle(), backpart_all.data(), (int)backpart_all.size())); }
private:
	BeepComm *h_ = nullptr;
};

/// PMCMC::obs_prob (pmcmc.cc:180-207) + bootstrap (:211-268) + Mpi::end_sec_swap (mpi.cc:215-327): log marginal
/// likelihood estimate of the bootstrap particle filter for one parameter vector shared by all particles.
/// comm == nullptr: all particles on this GPU.  Otherwise `particles` holds this rank's shard; the weights are
/// all-gathered, every rank takes the identical resampling decision, only the distinct ancestors that live on
/// another rank travel; the estimate is identical on every rank.  `ancestry` (optional) receives the ancestor
/// vector after every section, the argument of beep_particle_sample (Mpi::particle_sample, mpi.cc:331-389).
inline double particle_filter(Engine &eng, BatchState &particles, const std::vector<double> &theta_all, uint64_t seed, double invT = 1.0,
                              Comm *comm = nullptr, std::vector<uint32_t> *ancestry = nullptr)
{
	particles.set_param(theta_all);
	const int N = particles.size(), nranks = comm ? comm->size() : 1, rank = comm ? comm->rank() : 0;
	double loglik = 0;
	std::vector<uint32_t> back((size_t)N * nranks);
	std::vector<uint64_t> ids(N);
	for (int i = 0; i < N; i++) ids[i] = (uint64_t)rank * N + i;          // stream id = global particle index
	if (ancestry) ancestry->clear();
	for (int s = 0; s < eng.nsection(); s++) {
		int ti, tf;
		eng.section_range(s, ti, tf);
		if (s > 0) {
			if (comm && nranks > 1) comm->end_sec_swap(particles, back);
			else particles.permute(back, BEEP_PERMUTE_SECTION_STATE);
		}
		particles.simulate(ti, tf, seed, s == 0 ? ids.data() : nullptr);
		std::vector<double> L = particles.obs_loglik(s);
		if (comm && nranks > 1) L = comm->allgather(L);
		double inc;
		check(beep_bootstrap(L.data(), (int)L.size(), invT, seed, s, &inc, back.data()));
		loglik += inc;
		if (ancestry) ancestry->insert(ancestry->end(), back.begin(), back.end());
	}
	return loglik;
}

} // namespace beep
#endif
