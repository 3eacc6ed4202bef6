// beep_b200_main.cc -- the reference's host code running on the B200 engine.
//
// Built by host/ref_adapter/build.sh from the UNMODIFIED reference objects (everything of src/ except main.o:
// TOML/CSV inputs, Details, Data, Model with its priors and bounds, MVN proposals, utils' generators) compiled
// where they lie, plus this file, linked against beep_b200/lib/libbeep_b200.so.  It starts exactly like the
// reference's main (main.cc:100-158: Inputs, Details, Mpi, Data, Model, seeding, ObservationModel), flattens the
// model once (export_desc.hh -> BeepModelDesc) and then runs the inference driver with the BATCHED engine calls
// in place of the per-particle State calls:
//
//   mode=abcsmc   ABCSMC::run (abcsmc.cc:26-131) -- proposals from the reference's own MVN::propose / particle
//                 sampler / Model::inbounds, a whole batch of them through beep_set_params -> beep_simulate ->
//                 beep_obs_loglik (State::simulate(param), abcsmc.cc:46,71), the importance weights of the batch
//                 by beep_mvn_kernel_sums (calculate_particle_weight, :260-271), acceptance against the last
//                 cut-off (:76-79), remove_small_weights (:136-155), terminate on N_eff (:190-206),
//                 implement_cutoff_frac (:159-186), exchange_samples over the ranks (mpi.cc:169-211) by
//                 beep_allgather_scalar, normalise_particle_weights (:274-283).
//   mode=pmcmc    the reference's OWN PMCMC driver -- PMCMC::run (pmcmc.cc:33-110) with its burn-in, its ParamProp
//                 proposals and adaptation, its outputs -- object code unchanged, with the two members that touch the
//                 latent state replaced AT LINK TIME (build.sh weakens the two symbols in copies of pmcmc.o / mpi.o):
//                   PMCMC::obs_prob       pmcmc.cc:180-207  -> the bootstrap particle filter on the device (sections
//                                         simulated by beep_simulate, section likelihoods, beep_bootstrap, ancestors
//                                         moved by beep_permute / beep_end_sec_swap)
//                   Mpi::particle_sample  mpi.cc:331-389    -> beep_particle_sample (backward ancestry walk on the device)
//                 so every likelihood estimate of the Metropolis-Hastings chain comes from the engine and everything
//                 else -- acceptance, proposal tuning, trace and posterior files -- is the reference's.
//
//   BEEP_RUN=filter with any mode runs the bare filter (BEEP_NPARTICLE particles, BEEP_FILTER_REPS times, sharded over
//                 the ranks) -- also on models the reference refuses to load in mode=pmcmc (England_AS: observation
//                 windows cross its section boundaries).
//
// One process per GPU (RANK / WORLD_SIZE / LOCAL_RANK as torchrun or mpirun export them, or rank= nranks= on
// the command line); the ncclUniqueId travels through a file (BEEP_NCCL_ID_FILE, default /tmp/beep_nccl_id_<port>).
// Output: <outputdir>/b200_generations.csv and b200_posterior_parameters.csv, one line per generation on stdout.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

#define private public
#define protected public
#include "inputs.hh"
#include "details.hh"
#include "data.hh"
#include "model.hh"
#include "obsmodel.hh"
#include "mpi.hh"
#include "state.hh"
#include "output.hh"
#include "pmcmc.hh"
#include "ml.hh"
#include "mvn.hh"
#include "utils.hh"
#include "timers.hh"
#undef private
#undef protected

#include "export_desc.hh"
#include "beep_b200.hpp"

using namespace std;

static double now_s(){ return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); }
static int env_int(const char *k, int def){ const char *v = getenv(k); return v ? atoi(v) : def; }

struct Ranks {
	int rank = 0, nranks = 1, device = 0;
	BeepComm *comm = nullptr;
	vector <double> allgather(const vector <double> &loc) const
	{
		vector <double> all(loc.size()*nranks);
		if(loc.size() == 0) return all;
		if(nranks == 1) return loc;
		beep::check(beep_allgather_scalar(comm,loc.data(),(int)loc.size(),all.data()));
		return all;
	}
	/// mpi.combine (mpi.cc:641-676): variable-length vectors of every rank, concatenated in rank order
	vector <double> combine(const vector <double> &loc) const
	{
		if(nranks == 1) return loc;
		auto cnt = allgather(vector <double>(1,(double)loc.size()));
		size_t mx = 0; for(auto c : cnt) mx = max(mx,(size_t)c);
		if(mx == 0) return vector <double>();
		vector <double> pad(mx,0.0); copy(loc.begin(),loc.end(),pad.begin());
		auto all = allgather(pad);
		vector <double> out;
		for(int r = 0; r < nranks; r++) out.insert(out.end(),all.begin()+r*mx,all.begin()+r*mx+(size_t)cnt[r]);
		return out;
	}
	double sum(double v) const { double s = 0; for(auto x : allgather(vector <double>(1,v))) s += x; return s; }
	double maxv(double v) const { double s = v; for(auto x : allgather(vector <double>(1,v))) s = max(s,x); return s; }
};

static void setup_comm(Ranks &rk, beep::Engine &eng)
{
	if(rk.nranks == 1) return;
	string idfile = getenv("BEEP_NCCL_ID_FILE") ? getenv("BEEP_NCCL_ID_FILE") : string("/tmp/beep_nccl_id_")+(getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0");
	char id[128];
	if(rk.rank == 0){
		beep::check(beep_comm_unique_id(id));
		FILE *f = fopen((idfile+".tmp").c_str(),"wb"); if(!f){ perror(idfile.c_str()); exit(1);}
		fwrite(id,1,128,f); fclose(f);
		rename((idfile+".tmp").c_str(),idfile.c_str());
	}
	else{
		for(int spin = 0;; spin++){
			FILE *f = fopen(idfile.c_str(),"rb");
			if(f){ auto n = fread(id,1,128,f); fclose(f); if(n == 128) break;}
			if(spin > 60000){ cerr << "timed out waiting for " << idfile << endl; exit(1);}
			this_thread::sleep_for(chrono::milliseconds(10));
		}
	}
	beep::check(beep_comm_create(eng.handle(),id,rk.nranks,rk.rank,&rk.comm));
	rk.sum(0);                                                           // everyone is here: the file can go
	if(rk.rank == 0) remove(idfile.c_str());
}

// ---------------------------------------------------------------------------------------------------------
// ABC-SMC, abcsmc.cc:26-131, a batch of proposals at a time
// ---------------------------------------------------------------------------------------------------------
struct AbcSmc {
	const Details &details; const Model &model; Inputs &inputs; beep::Engine &eng; const Ranks &rk;
	unsigned int Ntot, Ntot_final, G; double cutoff_final, cpu_time, cutoff_frac, cutoff_frac_init, propsize;
	unsigned int B;                                                      // proposals per rank and round
	unsigned long long nstream = 0;                                      // simulations so far on this rank: the Philox stream ids
	vector <Generation> generation;
	vector <double> wsum;

	AbcSmc(const Details &details, const Model &model, Inputs &inputs, beep::Engine &eng, const Ranks &rk) : details(details), model(model), inputs(inputs), eng(eng), rk(rk)
	{
		inputs.find_nsample(Ntot,1);                                       // abcsmc.cc:17-21
		inputs.find_nsample_final(Ntot_final,Ntot);
		inputs.find_generation_or_cutoff_final_or_cpu_time(G,cutoff_final,cpu_time);
		inputs.find_cutoff_frac(cutoff_frac,cutoff_frac_init);
		inputs.find_propsize(propsize);
		auto co = min(cutoff_frac,cutoff_frac_init);
		B = (unsigned int)env_int("BEEP_BATCH",(int)ceil(Ntot/co/rk.nranks));
		B = max(16u,min(B,1u<<20));
	}

	unsigned int effective_particle_number(double sum, double sum2) const { return sum2 > 0 ? (unsigned int)(TINY+sum*sum/sum2) : 0; }   // abcsmc.cc:251-256

	bool terminate(const Generation &gen) const                          // abcsmc.cc:190-206
	{
		double s = 0, s2 = 0; for(auto w : gen.w){ s += w; s2 += w*w;}
		s = rk.sum(s); s2 = rk.sum(s2);
		auto Neff = effective_particle_number(s,s2);
		auto N = Ntot; if(gen.num == G-1) N = Ntot_final;
		auto co = cutoff_frac; if(gen.num == 0) co = cutoff_frac_init;
		return Neff >= N/co;
	}

	/// keeps the samples with keep[p] set (the reference erases one by one; same order, same survivors)
	void filter(Generation &gen, const vector <char> &keep) const
	{
		size_t k = 0;
		for(size_t p = 0; p < keep.size(); p++){
			if(!keep[p]) continue;
			if(k != p){ gen.param_samp[k] = std::move(gen.param_samp[p]); gen.w[k] = gen.w[p]; gen.EF_datatable[k] = std::move(gen.EF_datatable[p]);}
			k++;
		}
		gen.param_samp.resize(k); gen.w.resize(k); gen.EF_datatable.resize(k);
	}

	void remove_small_weights(Generation &gen) const                     // abcsmc.cc:136-155
	{
		double wmax = 0; for(auto w : gen.w) wmax = max(wmax,w);
		wmax = rk.maxv(wmax);
		vector <char> keep(gen.w.size()); for(size_t p = 0; p < keep.size(); p++) keep[p] = !(gen.w[p] < 0.000001*wmax);
		filter(gen,keep);
	}

	void implement_cutoff_frac(Generation &gen) const                    // abcsmc.cc:159-186
	{
		vector <double> EF_samp; for(const auto &ps : gen.param_samp) EF_samp.push_back(ps.EF);
		auto EFtot = rk.combine(EF_samp);
		sort(EFtot.begin(),EFtot.end());
		auto co = cutoff_frac; if(gen.num == 0) co = cutoff_frac_init;
		double EFcut = EFtot[(unsigned int)(co*EFtot.size())];
		if(cutoff_final != UNSET && EFcut < cutoff_final) EFcut = cutoff_final;
		vector <char> keep(gen.w.size()); for(size_t p = 0; p < keep.size(); p++) keep[p] = !(gen.param_samp[p].EF >= EFcut);
		filter(gen,keep);
		gen.EFcut = EFcut;
	}

	void exchange_samples(Generation &gen) const                         // mpi.cc:169-211
	{
		if(rk.nranks == 1) return;
		auto nparam = model.param.size(), ntab = gen.EF_datatable.size() ? gen.EF_datatable[0].size() : 0;
		vector <double> flat;
		for(auto i = 0u; i < gen.param_samp.size(); i++){
			flat.push_back(gen.param_samp[i].EF); flat.push_back(gen.w[i]);
			for(auto v : gen.param_samp[i].paramval) flat.push_back(v);
			for(auto v : gen.EF_datatable[i]) flat.push_back(v);
		}
		ntab = (size_t)rk.maxv((double)ntab);
		auto all = rk.combine(flat);
		auto rec = 2+nparam+ntab;
		gen.param_samp.clear(); gen.w.clear(); gen.EF_datatable.clear();
		for(size_t o = 0; o+rec <= all.size(); o += rec){
			ParamSample ps; ps.run = UNSET; ps.EF = all[o]; ps.paramval.assign(all.begin()+o+2,all.begin()+o+2+nparam);
			gen.param_samp.push_back(ps); gen.w.push_back(all[o+1]);
			gen.EF_datatable.push_back(vector <double>(all.begin()+o+2+nparam,all.begin()+o+rec));
		}
	}

	unsigned int particle_sampler() const                                // abcsmc.cc:238-247
	{
		auto NN = wsum.size();
		double z = ran()*wsum[NN-1];
		auto p = 0u; while(p < NN && z > wsum[p]) p++;
		if(p == NN) emsgEC("ABCSMC",2);
		return p;
	}

	void run()
	{
		const auto nparam = (unsigned int)model.param.size();
		const int ntable = eng.ntable();
		beep::BatchState batch(eng,(int)B,0);
		MVN mvn("ABCSMC",model.param_not_fixed,propsize,ALL_PARAM,MULTIPLE);
		vector <int> var; for(auto v : model.param_not_fixed) var.push_back((int)v);
		vector <double> theta((size_t)B*nparam), L(B), EFtab((size_t)B*max(ntable,1));
		vector <uint64_t> ids(B);
		const uint64_t seed = (uint64_t)inputs.find_integer("seed",0);
		const bool verbose = env_int("BEEP_VERBOSE",0) != 0;
		string dir = details.output_directory;
		mkdir(dir.c_str(),0755);
		ofstream gcsv;
		if(rk.rank == 0){ gcsv.open((dir+"/b200_generations.csv").c_str()); gcsv << "generation,EFcut,samples,simulations,acceptance,seconds,simulations_per_second" << endl;}

		const double t_start = now_s();
		auto g = 0u;
		do{
			const double t0 = now_s();
			Generation gen; gen.num = g;
			unsigned long long nsim = 0, ntr = 0, nac = 0;
			double wtot = 0, wcut = 0;
			vector <double> plast, wlast, inv_M;
			if(g > 0){
				Generation &gen_last = generation[g-1];
				mvn.setup(gen_last.param_samp,gen_last.w);                     // abcsmc.cc:55
				wsum.resize(gen_last.w.size()); double s = 0; for(auto i = 0u; i < wsum.size(); i++){ s += gen_last.w[i]; wsum[i] = s;}   // :223-234
				for(const auto &ps : gen_last.param_samp) plast.insert(plast.end(),ps.paramval.begin(),ps.paramval.end());
				wlast = gen_last.w;
				for(const auto &row : mvn.inv_M) inv_M.insert(inv_M.end(),row.begin(),row.end());
			}
			do{
				// one round: up to B parameter vectors (prior samples, or the in-bounds ones of at most 64 B MVN proposals
				// around sampled particles: an out-of-bounds proposal counts as tried and is dropped, abcsmc.cc:62-69);
				// slots that stay empty repeat the round's first vector and are ignored afterwards
				unsigned int nfill = 0;
				for(unsigned long long tries = 0; nfill < B && tries < 64ull*B; tries++){
					vector <double> param;
					if(g == 0) param = model.sample_from_prior();                // abcsmc.cc:44
					else{
						ntr++;
						param = mvn.propose(generation[g-1].param_samp[particle_sampler()].paramval);
						if(model.inbounds(param) == false) continue;
					}
					copy(param.begin(),param.end(),theta.begin()+(size_t)nfill*nparam);
					ids[nfill] = ((uint64_t)rk.rank << 40) + nstream++;
					nfill++;
				}
				if(verbose) fprintf(stderr,"[rank %d] generation %u: %u of %u slots filled, %llu proposals tried so far, %zu samples kept\n",rk.rank,g,nfill,B,ntr,gen.param_samp.size());
				if(rk.sum((double)nfill) == 0) emsgroot("ABC-SMC: no proposal fell inside the prior bounds (64 tries per slot): reduce 'propsize'");
				if(nfill == 0) copy(generation[g-1].param_samp[0].paramval.begin(),generation[g-1].param_samp[0].paramval.end(),theta.begin());
				for(auto i = max(nfill,1u); i < B; i++){ copy(theta.begin(),theta.begin()+nparam,theta.begin()+(size_t)i*nparam); ids[i] = ids[0];}
				batch.set_param(theta);                                        // State::simulate(param) = set_param + simulate + set_EF, state.cc:582-592
				batch.simulate(0,details.ndivision,seed,ids.data());
				beep::check(beep_obs_loglik(batch.handle(),-1,L.data(),ntable > 0 ? EFtab.data() : nullptr));
				nsim += nfill;
				vector <double> w(B,1.0);
				if(g > 0){
					auto sums = eng.mvn_kernel_sums(var,inv_M,mvn.size,(int)nparam,theta,plast,wlast);      // abcsmc.cc:260-271
					for(auto i = 0u; i < B; i++){
						vector <double> param(theta.begin()+(size_t)i*nparam,theta.begin()+(size_t)(i+1)*nparam);
						w[i] = exp(model.prior(param))/sums[i];
					}
				}
				for(auto i = 0u; i < nfill; i++){
					const double EF = -2*L[i];                                   // state.cc:95-98
					if(g > 0){ wtot += w[i]; if(!(EF < generation[g-1].EFcut)) continue; nac++; wcut += w[i];}
					ParamSample ps; ps.run = UNSET; ps.EF = EF; ps.paramval.assign(theta.begin()+(size_t)i*nparam,theta.begin()+(size_t)(i+1)*nparam);
					gen.param_samp.push_back(ps); gen.w.push_back(w[i]);
					gen.EF_datatable.push_back(vector <double>(EFtab.begin()+(size_t)i*ntable,EFtab.begin()+(size_t)(i+1)*ntable));
				}
				if(g > 0) remove_small_weights(gen);
			}while(!terminate(gen));
			if(g > 0) gen.model_evidence.push_back(log(rk.sum(wcut)/rk.sum(wtot)));
			const double acrate = g == 0 ? 1.0 : rk.sum((double)nac)/rk.sum((double)ntr);
			implement_cutoff_frac(gen);
			exchange_samples(gen);
			{ double ws = 0; for(auto w : gen.w) ws += w; ws /= gen.w.size(); for(auto &w : gen.w) w /= ws;}      // abcsmc.cc:274-283
			generation.push_back(gen);
			const double dt = now_s()-t0, nsim_all = rk.sum((double)nsim);
			if(rk.rank == 0){
				printf("generation %u: EFcut=%.6g samples=%zu simulations=%.0f acceptance=%.4f seconds=%.4f simulations/s=%.1f\n",g,gen.EFcut,gen.param_samp.size(),nsim_all,acrate,dt,nsim_all/dt);
				fflush(stdout);
				gcsv << g << "," << gen.EFcut << "," << gen.param_samp.size() << "," << nsim_all << "," << acrate << "," << dt << "," << nsim_all/dt << endl;
			}
			g++;
		}while(!(generation[g-1].EFcut == cutoff_final || g == G-1 || rk.maxv((now_s()-t_start)/60.0) > cpu_time));      // terminate_generation, abcsmc.cc:210-222

		if(rk.rank == 0){
			const auto &gen = generation.back();
			ofstream p((dir+"/b200_posterior_parameters.csv").c_str());
			p << "EF,weight"; for(const auto &pa : model.param) p << "," << pa.name; p << endl;
			for(auto i = 0u; i < gen.param_samp.size(); i++){
				p << gen.param_samp[i].EF << "," << gen.w[i]; for(auto v : gen.param_samp[i].paramval) p << "," << v; p << endl;
			}
		}
	}
};

// ---------------------------------------------------------------------------------------------------------
// bootstrap particle filter of PMCMC, pmcmc.cc:180-268, over the ranks
// ---------------------------------------------------------------------------------------------------------
static void run_filter(const Details &details, const Model &model, Inputs &inputs, const ObservationModel &obsmodel, beep::Engine &eng, const Ranks &rk)
{
	unsigned int Ntot = (unsigned int)env_int("BEEP_NPARTICLE",inputs.find_integer("nparticle",UNSET));
	if(Ntot == UNSET) emsgroot("'nparticle' must be set");
	if(Ntot%rk.nranks != 0) emsgroot("'nparticle' must be a multiple of the number of ranks");
	const int N = Ntot/rk.nranks, nsec = eng.nsection(), reps = env_int("BEEP_FILTER_REPS",3);
	const auto nparam = model.param.size();
	vector <double> param = model.simulated_values();
	for(auto v : param) if(v == UNSET){ param = model.sample_from_prior(); break;}
	param = rk.allgather(param); param.resize(nparam);                    // every rank runs rank 0's parameters
	vector <double> theta((size_t)N*nparam);
	for(int i = 0; i < N; i++) copy(param.begin(),param.end(),theta.begin()+(size_t)i*nparam);
	beep::BatchState part(eng,N,BEEP_KEEP_TRANSNUM);
	vector <uint64_t> ids(N); for(int i = 0; i < N; i++) ids[i] = (uint64_t)rk.rank*N+i;
	const uint64_t seed = (uint64_t)inputs.find_integer("seed",0);
	for(int rep = 0; rep < reps; rep++){
		const double t0 = now_s();
		double tswap = 0, tgather = 0;
		part.set_param(theta);
		double loglik = 0;
		vector <uint32_t> back((size_t)nsec*Ntot);
		for(int s = 0; s < nsec; s++){
			int ti, tf; eng.section_range(s,ti,tf);
			if(s > 0){
				const double t1 = now_s();
				if(rk.comm) beep::check(beep_end_sec_swap(rk.comm,part.handle(),&back[(size_t)(s-1)*Ntot],(int)Ntot));   // Mpi::end_sec_swap, pmcmc.cc:264
				else part.permute(vector <uint32_t>(back.begin()+(size_t)(s-1)*Ntot,back.begin()+(size_t)s*Ntot),BEEP_PERMUTE_SECTION_STATE);
				tswap += now_s()-t1;
			}
			part.simulate(ti,tf,seed+rep,s == 0 ? ids.data() : nullptr);
			auto L = part.obs_loglik(s);
			const double t2 = now_s();
			auto Lall = rk.allgather(L);                                     // mpi.gather(L), pmcmc.cc:219
			tgather += now_s()-t2;
			double inc;
			beep::check(beep_bootstrap(Lall.data(),(int)Ntot,1.0,seed+rep,s,&inc,&back[(size_t)s*Ntot]));
			loglik += inc;
		}
		vector <double> th(nparam), tn((size_t)details.ndivision*eng.cells_per_step());
		beep::check(beep_particle_sample(rk.comm,part.handle(),back.data(),nsec,(int)Ntot,th.data(),tn.data()));   // Mpi::particle_sample, pmcmc.cc:154
		const double dt = now_s()-t0;
		if(rk.rank == 0){
			double ev = 0; for(auto v : tn) ev += v;
			printf("filter %d: particles=%u sections=%d log-likelihood=%.10g seconds=%.4f swap_seconds=%.4f gather_seconds=%.4f sampled_events=%.0f\n",rep,Ntot,nsec,loglik,dt,tswap,tgather,ev);
			fflush(stdout);
		}
	}
	(void)obsmodel;
}

// ---------------------------------------------------------------------------------------------------------
// mode=pmcmc: the reference's PMCMC object on the engine.  The two definitions below replace the reference's own
// (weak in the linked copies of pmcmc.o / mpi.o, see build.sh); everything else of PMCMC runs as compiled.
// ---------------------------------------------------------------------------------------------------------
struct PmcmcEngine {
	beep::Engine *eng = nullptr;
	beep::BatchState *part = nullptr;
	const Ranks *rk = nullptr;
	const Details *details = nullptr;
	uint64_t seed = 0, ncall = 0;
	vector <double> theta;
	double seconds = 0;
};
static PmcmcEngine g_pm;

/// pmcmc.cc:180-207 + bootstrap :211-268 + Mpi::end_sec_swap mpi.cc:215-327
double PMCMC::obs_prob(vector <double> &paramv)
{
	const double t0 = now_s();
	auto &part = *g_pm.part;
	const int n = part.size(), nsec = g_pm.eng->nsection();
	const auto nparam = paramv.size();
	g_pm.theta.resize((size_t)n*nparam);
	for(int i = 0; i < n; i++) copy(paramv.begin(),paramv.end(),g_pm.theta.begin()+(size_t)i*nparam);
	part.set_param(g_pm.theta);
	const uint64_t seed = g_pm.seed+(++g_pm.ncall);                      // a fresh set of streams per likelihood estimate
	vector <uint64_t> ids(n); for(int i = 0; i < n; i++) ids[i] = (uint64_t)g_pm.rk->rank*n+i;
	auto obprob = 0.0;
	vector <uint32_t> back(Ntot);
	for(int s = 0; s < nsec; s++){
		int ti, tf; g_pm.eng->section_range(s,ti,tf);
		if(s > 0){
			if(g_pm.rk->comm) beep::check(beep_end_sec_swap(g_pm.rk->comm,part.handle(),back.data(),(int)Ntot));
			else part.permute(back,BEEP_PERMUTE_SECTION_STATE);
		}
		part.simulate(ti,tf,seed,s == 0 ? ids.data() : nullptr);
		auto Lall = g_pm.rk->allgather(part.obs_loglik(s));                // mpi.gather(L), pmcmc.cc:219
		double inc;
		beep::check(beep_bootstrap(Lall.data(),(int)Ntot,invT,seed,s,&inc,back.data()));
		obprob += inc;
		for(auto p = 0u; p < Ntot; p++) backpart[s+1][p] = back[p];         // the reference's layout: row sec+1 = ancestors after section sec
		if(std::isnan(obprob) || std::isinf(obprob)) emsgEC("PMCMC",1);
	}
	g_pm.seconds += now_s()-t0;
	return obprob;
}

/// mpi.cc:331-389: one filtered trajectory through the ancestry of the LAST obs_prob
Particle Mpi::particle_sample(const unsigned int ru, const vector < vector <unsigned int> > &backpart, vector <State> &particle, const ObservationModel &obsmodel)
{
	const int nsec = (int)obsmodel.nsection, ntot = (int)backpart[0].size();
	vector <uint32_t> back((size_t)nsec*ntot);
	for(int s = 0; s < nsec; s++) for(int p = 0; p < ntot; p++) back[(size_t)s*ntot+p] = backpart[s+1][p];
	const auto &det = *g_pm.details;
	Particle pa;
	pa.run = ru; pa.EF = 0;
	pa.paramval.resize(g_pm.theta.size()/g_pm.part->size());
	vector <double> tn((size_t)det.ndivision*g_pm.eng->cells_per_step());
	beep::check(beep_particle_sample(g_pm.rk->comm,g_pm.part->handle(),back.data(),nsec,ntot,pa.paramval.data(),tn.data()));
	int A = 0, J = 0, D = 0;
	beep::check(beep_engine_dims(g_pm.eng->handle(),nullptr,&A,nullptr,&J,&D,nullptr,nullptr,nullptr));
	pa.transnum.resize(det.ndivision);
	size_t k = 0;
	for(auto sett = 0u; sett < det.ndivision; sett++){
		pa.transnum[sett].resize(A);
		for(int c = 0; c < A; c++){
			pa.transnum[sett][c].resize(J);
			for(int tr = 0; tr < J; tr++){ pa.transnum[sett][c][tr].assign(tn.begin()+k,tn.begin()+k+D); k += D;}
		}
	}
	(void)particle;
	return pa;
}

static void run_pmcmc(const Details &details, const Data &data, const Model &model, Inputs &inputs, const ObservationModel &obsmodel, Mpi &mpi, beep::Engine &eng, const Ranks &rk)
{
	if(rk.nranks != 1) emsgroot("mode=pmcmc through the adapter runs the chain on one rank (the filter itself shards: BEEP_RUN=filter)");
	Output output(details,data,model,inputs,obsmodel,mpi);
	PMCMC pmcmc(details,data,model,inputs,output,obsmodel,mpi);          // the reference's object: nparticle, nsample, nburnin, invT from the TOML / command line
	beep::BatchState part(eng,(int)pmcmc.N,BEEP_KEEP_TRANSNUM);
	g_pm.eng = &eng; g_pm.part = &part; g_pm.rk = &rk; g_pm.details = &details;
	g_pm.seed = (uint64_t)inputs.find_integer("seed",0)*1000003ull;
	const double t0 = now_s();
	pmcmc.run();
	cout << "pmcmc on the engine: " << g_pm.ncall << " particle filters of " << pmcmc.Ntot << " particles, " << g_pm.seconds << " s in the filter of " << now_s()-t0 << " s" << endl;
}

// ---------------------------------------------------------------------------------------------------------
// mode=map: the reference's ML object (CMA-ES, ml.cc:150-385) on the engine.  ML::generate_samples (ml.cc:389-413)
// -- the loop that evaluates State::likelihood_approx for every sample of a generation -- is replaced at link time
// (weak in the linked copy of ml.o, see build.sh) by ONE beep_likelihood_approx call for the whole generation;
// sampling (ML::sample_param), priors, the CMA-ES update, the Hessian and the final posterior samples run as compiled.
// ---------------------------------------------------------------------------------------------------------
struct MlEngine {
	beep::Engine *eng = nullptr;
	beep::BatchState *part = nullptr;
	long ncall = 0, nlik = 0;
	double seconds = 0;
	vector <double> theta;
};
static MlEngine g_ml;

void ML::generate_samples(vector <ParamSample> &ps_per_core, const vector <double> &mean, const vector < vector <double> > &C, const double sigma, Generation &gen)
{
	auto psamp = sample_param(mean,C,sigma,ps_per_core.size());          // the reference's own sampler (ml.cc:417-441): same ran() stream
	if(details.stochastic != true) emsgroot("mode=map through the adapter evaluates the approximate likelihood of stochastic models");
	const auto n = ps_per_core.size(), nparam = model.param.size();
	if((int)n != g_ml.part->size()) emsgroot("generation size differs from the engine batch");
	const double t0 = now_s();
	g_ml.theta.resize(n*nparam);
	for(auto i = 0u; i < n; i++) copy(psamp[i].begin(),psamp[i].end(),g_ml.theta.begin()+(size_t)i*nparam);
	g_ml.part->set_param(g_ml.theta);
	auto L = g_ml.part->likelihood_approx(invT);
	g_ml.seconds += now_s()-t0; g_ml.ncall++; g_ml.nlik += n;
	for(auto i = 0u; i < n; i++){
		ps_per_core[i].paramval = psamp[i];
		auto Pr = model.prior(psamp[i]);
		ps_per_core[i].EF = -2*(L[i] + Pr);
		if(ps_per_core[i].EF < 0) emsgEC("ML",44);
		state.paramval = psamp[i];                                         // what set_param leaves behind; EF and the trajectory are diagnostics only
		gen.param_samp.push_back(state.create_param_sample(0));
		gen.EF_datatable.push_back(vector <double>(data.datatable.size(),0.0));
	}
}

static void run_map(const Details &details, const Data &data, const Model &model, Inputs &inputs, const ObservationModel &obsmodel, Mpi &mpi, beep::Engine &eng, const Ranks &rk)
{
	if(rk.nranks != 1) emsgroot("mode=map through the adapter runs on one rank");
	if(!beep_engine_has_approx(eng.handle())) emsgroot("the engine refuses this model for the approximate likelihood");
	Output output(details,data,model,inputs,obsmodel,mpi);
	ML ml(details,data,model,inputs,output,obsmodel,mpi);                // the reference's object: algorithm, nparticle, ngeneration, invT from the TOML / command line
	beep::BatchState part(eng,(int)ml.npart_per_core);
	g_ml.eng = &eng; g_ml.part = &part;
	const double t0 = now_s();
	ml.run();
	cout << "map on the engine: " << g_ml.ncall << " generations of " << ml.npart_per_core << " approximate likelihoods, " << g_ml.seconds << " s on the engine of " << now_s()-t0 << " s" << endl;
}

int main(int argc, char **argv)
{
	timersinit();
	Ranks rk;
	rk.rank = env_int("RANK",env_int("OMPI_COMM_WORLD_RANK",0)); rk.nranks = env_int("WORLD_SIZE",env_int("OMPI_COMM_WORLD_SIZE",1));
	rk.device = env_int("LOCAL_RANK",rk.rank);

	Inputs inputs(argc,argv);                                            // main.cc:123-151, unchanged order
	Details details(inputs);
	Mpi mpi(details);
	Data data(inputs,details,mpi);
	Model model(inputs,details,data,mpi);
	auto seed = inputs.find_integer("seed",0);
	sran(rk.rank*10000+seed+100);                                        // main.cc:153-158 with the rank in place of mpi.core
	const bool filter_only = getenv("BEEP_RUN") && string(getenv("BEEP_RUN")) == "filter";
	if(details.mode != PMCMC_INF){
		// the sections of the particle filter by the rule of details.cc:110-121 (ten sections on the coarsest data
		// time step), whatever the mode: observations that cross a section boundary belong to no section
		// (obsmodel.cc:326-348) -- England_AS as shipped cannot be loaded in mode=pmcmc for that reason
		auto max_dt = inputs.find_maximum_timestep();
		details.pmcmc_obs_period = details.division_per_time*max_dt*int(details.period/(max_dt*10.0));
		if(details.pmcmc_obs_period == 0) details.pmcmc_obs_period = details.ndivision;
	}
	ObservationModel obsmodel(details,data,model);

	DescMemory dm;
	beep_export_desc(dm,details,data,model);
	BeepModelDesc desc = dm.desc();
	try{
		beep::Engine eng(desc,rk.device);
		setup_comm(rk,eng);
		if(rk.rank == 0) cout << "BEEP on the B200 engine: " << rk.nranks << " rank(s), T=" << desc.T << " A=" << desc.A << " D=" << desc.D << " C=" << desc.C << " J=" << desc.J << " nparam=" << desc.nparam << " nobs=" << desc.nobs << endl;
		if(filter_only) run_filter(details,model,inputs,obsmodel,eng,rk);
		else switch(details.mode){
		case ABC_SMC: { AbcSmc abc(details,model,inputs,eng,rk); abc.run(); } break;
		case PMCMC_INF: run_pmcmc(details,data,model,inputs,obsmodel,mpi,eng,rk); break;
		case ML_INF: run_map(details,data,model,inputs,obsmodel,mpi,eng,rk); break;
		default: emsgroot("beep_b200_main runs mode=abcsmc, mode=pmcmc, mode=map and BEEP_RUN=filter"); break;
		}
		if(rk.comm) beep_comm_destroy(rk.comm);
	}
	catch(const beep::Error &e){ cerr << "engine error " << e.code << ": " << e.what() << endl; return 1; }
	return 0;
}
