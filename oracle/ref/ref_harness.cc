// ref_harness.cc -- TEST INFRASTRUCTURE (never shipped, never on the product path).
//
// Links against the UNMODIFIED reference objects compiled where they lie under
// /root/reference (oracle/ref/build_ref.sh) and does two jobs:
//
//  1. the reference-side ADAPTER: flattens the reference's Details / Data / Model
//     into the BeepModelDesc arrays of include/beep_model_fields.def (this is the
//     code a BEEP maintainer would add next to main.cc:125-162);
//  2. the GOLDEN-VECTOR generator: runs the reference's own State::set_param,
//     State::simulate, ObservationModel::calculate*, State::likelihood and
//     Mbp::mbp and records their outputs, so that oracle/ (the CPU restatement)
//     and the CUDA engine can be pinned against the real thing.
//
// For the stochastic runs the reference's poisson_sample / binomial_sample
// (utils.cc:148-153,174-180) are replaced AT LINK TIME (utils.cc is compiled with
// -Dpoisson_sample=beepref_poisson_sample ..., sources untouched) by the
// counter-based samplers of include/beep_rng.h consuming ONE sequential stream, so
// that integer event counts can be compared bit for bit.
//
// Output: a "BPK1" container (named little-endian arrays) read by
// tests/golden/make_golden.py.

#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <iostream>
#include <fstream>
#include <sstream>
#include <map>
#include <random>
#include <algorithm>

// The hot-path seam has private members we must read (Model::inf_state,
// Mbp::mbp/initial/propose/simu_or_mbp, ObservationModel::split_observations);
// access specifiers do not change layout, so open them for this TU only.
#define private public
#define protected public
#include "inputs.hh"
#include "details.hh"
#include "data.hh"
#include "model.hh"
#include "obsmodel.hh"
#include "state.hh"
#include "output.hh"
#include "mpi.hh"
#include "mbp.hh"
#include "utils.hh"
#undef private
#undef protected

#include "beep_rng.h"

using namespace std;

// ---------------------------------------------------------------------------
// link-time replacement of the reference samplers
// ---------------------------------------------------------------------------
int beepref_poisson_sample(const double lam);                       // original, utils.cc:148
unsigned int beepref_binomial_sample(const double p, const unsigned int n); // original, utils.cc:174

// The reference warm-starts its power iteration (matrix.cc:1136-1220) from a GLOBAL vector that
// is first filled with ran() and then carried across every call, so State::set_param depends on
// the call history at the 1e-13 level (its convergence tolerance).  For reproducible golden
// vectors the harness resets that vector to the flat start before each top-level call.
extern vector <double> initial_guess;
static unsigned int g_ngm_size = 0;
static void reset_guess(){ initial_guess.assign(g_ngm_size,1.0/g_ngm_size); }

static bool g_inject = false;
static BeepStream g_stream;
static long g_ndraw = 0;

int poisson_sample(const double lam)
{
	if(!g_inject) return beepref_poisson_sample(lam);
	g_ndraw++;
	return beep_poisson(&g_stream,lam);
}

unsigned int binomial_sample(const double p, const unsigned int n)
{
	if(!g_inject) return beepref_binomial_sample(p,n);
	if(n == 0) return 0;                                               // utils.cc:176
	g_ndraw++;
	return (unsigned int)beep_binomial(&g_stream,p,(int)n);
}

static void inject_on(uint64_t seed, uint32_t id)
{
	g_inject = true;
	beep_stream_init(&g_stream,seed,id,0,0xFFFFFFFFu);
}
static void inject_off(){ g_inject = false; }

// ---------------------------------------------------------------------------
// BPK1 container
// ---------------------------------------------------------------------------
struct Pack {
	FILE *f; unsigned int n;
	Pack(const char *file){ f = fopen(file,"wb"); if(!f){ perror(file); exit(1);} n = 0; fwrite("BPK1",1,4,f); fwrite(&n,4,1,f); }
	void head(const string &name, unsigned int dtype, const vector <long> &dims){
		unsigned int len = name.size(); fwrite(&len,4,1,f); fwrite(name.data(),1,len,f);
		fwrite(&dtype,4,1,f); unsigned int nd = dims.size(); fwrite(&nd,4,1,f);
		for(auto d : dims){ long long dd = d; fwrite(&dd,8,1,f);}
		n++;
	}
	void i32(const string &name, const vector <int> &v, vector <long> dims = {}){
		if(dims.size() == 0) dims.push_back(v.size());
		long tot = 1; for(auto d : dims) tot *= d; if(tot != (long)v.size()){ cerr << "dims " << name << endl; exit(1);}
		head(name,0,dims); if(v.size()) fwrite(v.data(),4,v.size(),f);
	}
	void f64(const string &name, const vector <double> &v, vector <long> dims = {}){
		if(dims.size() == 0) dims.push_back(v.size());
		long tot = 1; for(auto d : dims) tot *= d; if(tot != (long)v.size()){ cerr << "dims " << name << endl; exit(1);}
		head(name,1,dims); if(v.size()) fwrite(v.data(),8,v.size(),f);
	}
	void scalar_i(const string &name, int v){ i32(name,vector<int>{v},{1}); }
	void scalar_d(const string &name, double v){ f64(name,vector<double>{v},{1}); }
	void close(){ fseek(f,4,SEEK_SET); fwrite(&n,4,1,f); fclose(f); }
};

static int u2i(unsigned int v){ return v == UNSET ? -1 : (int)v; }

static vector <double> flat4(const vector < vector < vector < vector <double> > > > &x, unsigned int n0)
{
	vector <double> v;
	for(auto a = 0u; a < n0; a++) for(const auto &b : x[a]) for(const auto &c : b) for(auto d : c) v.push_back(d);
	return v;
}
static vector <double> flat3(const vector < vector < vector <double> > > &x)
{
	vector <double> v;
	for(const auto &a : x) for(const auto &b : a) for(auto c : b) v.push_back(c);
	return v;
}
static vector <double> flat2(const vector < vector <double> > &x)
{
	vector <double> v;
	for(const auto &a : x) for(auto b : a) v.push_back(b);
	return v;
}

// ---------------------------------------------------------------------------
// 1. adapter: Details + Data + Model -> BeepModelDesc arrays
// ---------------------------------------------------------------------------
// the exporter itself is product code (the reference-side adapter): host/ref_adapter/export_desc.hh
#include "../../host/ref_adapter/export_desc.hh"
static void export_desc(Pack &pk, const Details &details, const Data &data, const Model &model){ beep_export_desc(pk,details,data,model); }

// ---------------------------------------------------------------------------
// 2. golden vectors
// ---------------------------------------------------------------------------
static void dump_derived(Pack &pk, const string &pre, const State &st)
{
	pk.f64(pre+"transrate",flat2(st.transrate));
	pk.f64(pre+"disc_spline",flat2(st.disc_spline));
	pk.f64(pre+"susceptibility",st.susceptibility);
	pk.f64(pre+"areafactor",flat2(st.areafactor));
	pk.f64(pre+"beta",flat3(st.beta));
	// Ntime is N0[a][b] g_a g_b (model.cc:1254-1285): record it at three steps only
	auto T = st.Ntime.size();
	vector <double> nt;
	for(auto s : {0ul,T/2,T-1}) for(const auto &r : st.Ntime[s]) for(auto v : r) nt.push_back(v);
	pk.f64(pre+"Ntime3",nt);
	pk.scalar_d(pre+"genT",st.genT);
}

static void dump_traj(Pack &pk, const string &pre, const State &st, const ObservationModel &obsmodel, unsigned int T, bool with_pop)
{
	pk.f64(pre+"transnum",flat4(st.transnum,T));
	pk.f64(pre+"transmean",flat4(st.transmean,T));
	if(with_pop){
		pk.f64(pre+"pop",flat4(st.pop,T));
		pk.f64(pre+"Imap",flat3(st.Imap));
		pk.f64(pre+"Idiag",flat3(st.Idiag));
	}
	pk.scalar_d(pre+"EF",st.EF);
	pk.f64(pre+"obs_value",obsmodel.get_obs_value(&st));
	pk.f64(pre+"EF_datatable",obsmodel.get_EF_datatable(&st));
	pk.scalar_d(pre+"Li",st.likelihood());
	if(getenv("BEEP_HARNESS_GRAPHS")){                                  // ObservationModel::get_graph_state, obsmodel.cc:113-170
		vector <double> gs; for(const auto &g : obsmodel.get_graph_state(&st)) for(auto v : g) gs.push_back(v);
		pk.f64(pre+"graph_state",gs);
	}
}

int main(int argc, char **argv)
{
	if(argc < 5){ cerr << "usage: ref_harness out.bpk ntheta seed inputfile=... mode=... [key=value ...]" << endl; return 1;}
	string outfile = argv[1];
	int ntheta = atoi(argv[2]);
	int seed = atoi(argv[3]);
	bool big = getenv("BEEP_HARNESS_BIG") != NULL;                     // England_AS-sized: skip bulky arrays

	timersinit();
	vector <char*> av; av.push_back(argv[0]); for(int i = 4; i < argc; i++) av.push_back(argv[i]);

	Inputs inputs((int)av.size(),av.data());
	Details details(inputs);
	Mpi mpi(details);
	Data data(inputs,details,mpi);
	Model model(inputs,details,data,mpi);
	sran(seed);
	if(details.mode != PMCMC_INF){                                      // details.cc:110-121 (same rule, any mode)
		auto max_dt = inputs.find_maximum_timestep();
		details.pmcmc_obs_period = details.division_per_time*max_dt*int(details.period/(max_dt*10.0));
		if(details.pmcmc_obs_period == 0) details.pmcmc_obs_period = details.ndivision;
	}
	if(getenv("BEEP_HARNESS_OBSPATCH") != NULL){
		// No example of the reference uses a plain-normal or negative-binomial observation model, censored
		// ("thresh") or unknown data (negbin cannot even be selected from a TOML file,
		// inputs.cc:455-460).  To pin ObservationModel::obs_prob's remaining branches (obsmodel.cc:245-305) the
		// loaded observations are re-labelled in memory, cycling through every (model, datum kind) pair; the
		// golden likelihoods are then the reference's own obs_prob on them.
		for(auto tb = 0u; tb < data.datatable.size(); tb++){               // threshold = the table's largest datum: the censored sums stay representable
			vector <double> vals; for(const auto &ob : data.obs) if(ob.datatable == tb && ob.value < UNKNOWN) vals.push_back(ob.value);
			sort(vals.begin(),vals.end());
			data.datatable[tb].threshold = vals.size() ? (unsigned int)(vals[vals.size()-1]+1) : 6;
		}
		for(auto i = 0u; i < data.obs.size(); i++){
			auto &ob = data.obs[i];
			switch(i%4){
			case 0: ob.obsmodel = NORMAL_OBSMODEL; if(!(ob.sd > 0)) ob.sd = 2.5; break;
			case 1: ob.obsmodel = NORMAL_PERCENT_OBSMODEL; if(!(ob.sd > 0)) ob.sd = 1.5; break;
			case 2: ob.obsmodel = POISSON_OBSMODEL; break;
			case 3: ob.obsmodel = NEGBINO_OBSMODEL; ob.shape = 3.5 + (i%5); break;
			}
			if((i/4)%3 == 1) ob.value = THRESH;
			if((i/4)%5 == 2) ob.value = UNKNOWN;
		}
	}
	ObservationModel obsmodel(details,data,model);
	if(details.mode != PMCMC_INF) obsmodel.split_observations();       // obsmodel.cc:326-348
	Output output(details,data,model,inputs,obsmodel,mpi);

	Pack pk(outfile.c_str());
	export_desc(pk,details,data,model);
	g_ngm_size = data.ndemocatpos_per_strain;

	auto T = details.ndivision;
	pk.scalar_i("meta.ntheta",ntheta);
	pk.scalar_i("meta.seed",seed);
	pk.scalar_i("meta.nsection",obsmodel.nsection);
	{ vector <int> ti, tf; for(auto s = 0u; s < obsmodel.nsection; s++){ ti.push_back(obsmodel.section_ti[s]); tf.push_back(obsmodel.section_tf[s]);}
		pk.i32("meta.section_ti",ti); pk.i32("meta.section_tf",tf);}

	// ---- parameter vectors: the TOML 'value's when they are complete, then prior samples
	vector < vector <double> > theta;
	{
		auto sv = model.simulated_values();
		bool ok = true; for(auto v : sv) if(v == UNSET || std::isnan(v)) ok = false;
		if(ok && model.inbounds(sv) == false && details.siminf != SIMULATE) ok = false;
		if(ok) theta.push_back(sv);
		auto tries = 0u;
		while((int)theta.size() < ntheta && tries < 1000){
			auto th = model.sample_from_prior(); tries++;
			if(model.inbounds(th) == true || details.siminf == SIMULATE) theta.push_back(th);
		}
		pk.f64("theta",flat2(theta),{(long)theta.size(),(long)model.param.size()});
		vector <double> pr; for(const auto &th : theta) pr.push_back(model.prior(th));
		pk.f64("prior",pr);
	}

	// ---- CPU-baseline mode (bench.py --impl reference): the reference's own State::simulate(param)
	// (set_param + tau-leap, state.cc:582-592) and ObservationModel::calculate (obsmodel.cc:22-35) with its
	// own std:: samplers, repeated for BEEP_HARNESS_BENCH seconds; nothing else is written
	if(getenv("BEEP_HARNESS_BENCH")){
		double budget = atof(getenv("BEEP_HARNESS_BENCH"));
		details.stochastic = true;
		State st(details,data,model,obsmodel);
		st.simulate(theta[0]);
		volatile double sink = obsmodel.calculate(&st);
		struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC,&t0);
		long reps = 0; double el = 0;
		do{
			st.simulate(theta[0]);
			sink = obsmodel.calculate(&st);
			reps++;
			clock_gettime(CLOCK_MONOTONIC,&t1); el = (t1.tv_sec-t0.tv_sec)+1e-9*(t1.tv_nsec-t0.tv_nsec);
		}while(el < budget);
		(void)sink;
		printf("BENCH reps=%ld seconds=%.6f T=%u A=%u D=%u\n",reps,el,(unsigned)T,(unsigned)data.narea,(unsigned)data.ndemocatpos);
		{ FILE *bf = fopen((outfile+".bench").c_str(),"w");                 // the reference prints its own progress on stdout: results go to a file
			if(bf){ fprintf(bf,"BENCH reps=%ld seconds=%.6f T=%u A=%u D=%u\n",reps,el,(unsigned)T,(unsigned)data.narea,(unsigned)data.ndemocatpos); fclose(bf);} }
		pk.close();
		return 0;
	}

	// ---- MVN-approximate likelihood of mode=map (ML::generate_samples, ml.cc:389-413): the reference's own
	// State::likelihood_approx (state_mvn_approx.cc:11-91) with ML's set-up (ml.cc:39-42), at the inverse temperature
	// given in BEEP_HARNESS_APPROX; mean populations and the variances it carries are recorded next to L
	if(getenv("BEEP_HARNESS_APPROX")){
		double invT = atof(getenv("BEEP_HARNESS_APPROX"));
		auto obs_slice = obsmodel.generate_obs_slice();
		State st(details,data,model,obsmodel);
		st.initialise_approx();
		vector <double> Ls;
		for(auto k = 0u; k < theta.size(); k++){
			reset_guess();
			auto L = st.likelihood_approx(theta[k],obs_slice,invT,DOUBLE);
			Ls.push_back(L);
			string pre = "th"+to_string(k)+".approx.";
			if(L != -LARGE){
				pk.f64(pre+"pop",flat4(st.pop,T+1));
				vector <double> var;
				for(auto sett = 0u; sett < T; sett++) for(auto c = 0u; c < st.ncomp_approx; c++) var.push_back(st.pop_covar[sett][c][c]);
				pk.f64(pre+"var",var);
			}
		}
		pk.f64("approx.L",Ls);
		pk.scalar_d("approx.invT",invT);
		pk.scalar_i("approx.ncomp",(int)st.ncomp_approx);
		pk.close();
		cout << "ref_harness: wrote " << outfile << " (approx, " << pk.n << " arrays)" << endl;
		return 0;
	}

	const bool stoch0 = details.stochastic;
	vector <Particle> stoch_part;

	for(auto k = 0u; k < theta.size(); k++){
		string pre = "th"+to_string(k)+".";
		State st(details,data,model,obsmodel);

		// (i) derived tables, State::set_param (state.cc:66-91)
		reset_guess();
		st.set_param(theta[k]);
		dump_derived(pk,pre,st);

		// (ii) deterministic trajectory (state.cc:627 with details.stochastic == false)
		details.stochastic = false;
		reset_guess();
		st.simulate(theta[k]);
		dump_traj(pk,pre+"det.",st,obsmodel,T,!big);
		{ vector <double> Ls; for(auto s = 0u; s < obsmodel.nsection; s++) Ls.push_back(obsmodel.calculate_section(&st,s)); pk.f64(pre+"det.L_section",Ls);}

		// (iii) stochastic trajectory, injected sequential stream
		details.stochastic = true;
		inject_on(seed,k); g_ndraw = 0;
		reset_guess();
		st.simulate(theta[k]);
		inject_off();
		pk.scalar_i(pre+"sto.ndraw",(int)g_ndraw);
		dump_traj(pk,pre+"sto.",st,obsmodel,T,false);
		{ vector <double> Ls; for(auto s = 0u; s < obsmodel.nsection; s++) Ls.push_back(obsmodel.calculate_section(&st,s)); pk.f64(pre+"sto.L_section",Ls);}
		stoch_part.push_back(st.create_particle(0));

		// (iv) sectioned simulation as PMCMC::obs_prob does (pmcmc.cc:190-199), same stream
		if(k == 0){
			inject_on(seed,1000+k);
			vector <double> Ls;
			for(auto s = 0u; s < obsmodel.nsection; s++){
				st.simulate(obsmodel.section_ti[s],obsmodel.section_tf[s]);
				Ls.push_back(obsmodel.calculate_section(&st,s));
			}
			inject_off();
			pk.f64(pre+"sec.transnum",flat4(st.transnum,T));
			pk.f64(pre+"sec.L_section",Ls);
		}
		details.stochastic = stoch0;
	}

	// ---- model-based proposals, Mbp::mbp (mbp.cc:252-330)
	if(theta.size() >= 2){
		details.stochastic = true;
		Mbp mbp(INVT,details,data,model,obsmodel,output,mpi);
		mbp.invT = 1; mbp.EFcut = UNSET;

		// (v) pure MBP from theta0's trajectory towards theta1
		reset_guess();
		mbp.initial->initialise_from_particle(stoch_part[0]);
		pk.f64("mbp.init.transmean",flat4(mbp.initial->transmean,T));
		pk.scalar_d("mbp.init.Li",mbp.initial->likelihood());
		inject_on(seed,2000); g_ndraw = 0;
		reset_guess();
		auto ok = mbp.mbp(theta[1],INF_DIF_UPDATE);
		inject_off();
		pk.scalar_i("mbp.a.status",ok == SUCCESS ? 1 : 0);
		pk.scalar_i("mbp.a.ndraw",(int)g_ndraw);
		if(ok == SUCCESS){
			mbp.propose->set_EF();
			pk.f64("mbp.a.transnum",flat4(mbp.propose->transnum,T));
			pk.f64("mbp.a.transmean",flat4(mbp.propose->transmean,T));
			pk.scalar_d("mbp.a.EF",mbp.propose->EF);
			pk.scalar_d("mbp.a.Li",mbp.propose->likelihood());
			// Mbp::get_al (mbp.cc:334-361) on this proposal in both modes: INVT at three temperatures, CUTOFF with the cut-off above and below EF_p
			mbp.initial->set_EF(); mbp.initial->set_Pr();
			vector <double> al, arg;
			for(auto it : {1.0,0.37,0.002}){ mbp.obsmodel_mode = INVT; mbp.invT = it; al.push_back(mbp.get_al()); arg.push_back(it);}
			for(auto dc : {3.0,-3.0}){ mbp.obsmodel_mode = CUTOFF; mbp.EFcut = mbp.propose->EF+dc; al.push_back(mbp.get_al()); arg.push_back(mbp.EFcut);}
			mbp.obsmodel_mode = INVT; mbp.invT = 1; mbp.EFcut = UNSET;
			pk.f64("mbp.al.value",al); pk.f64("mbp.al.arg",arg);
			pk.f64("mbp.al.state",vector <double>{mbp.initial->EF,mbp.propose->EF,mbp.initial->Pr,mbp.propose->Pr});
		}

		// (vi) mixed SIMU/MBP with the same parameters (mbp_slicetime / mbp_fixedtree pattern, mbp.cc:656-697)
		vector <int> mask;
		for(auto sett = 0u; sett < T; sett++){
			for(auto c = 0u; c < data.narea; c++){
				bool simu = (sett >= T/3 && sett < (2*T)/3 && (c%2 == 0));
				mbp.simu_or_mbp[sett][c] = simu ? SIMU : MBP;
				mask.push_back(simu ? 1 : 0);
			}
		}
		pk.i32("mbp.b.mask",mask,{(long)T,(long)data.narea});
		inject_on(seed,3000);
		reset_guess();
		ok = mbp.mbp(theta[0],INF_DIF_UPDATE);
		inject_off();
		pk.scalar_i("mbp.b.status",ok == SUCCESS ? 1 : 0);
		if(ok == SUCCESS){
			mbp.propose->set_EF();
			pk.f64("mbp.b.transnum",flat4(mbp.propose->transnum,T));
			pk.f64("mbp.b.transmean",flat4(mbp.propose->transmean,T));
			pk.scalar_d("mbp.b.EF",mbp.propose->EF);
		}
		mbp.simu_or_mbp_reset();

		// (vii) parameter proposal with fixed events (mbp.cc:460-491): transmean under theta1, transnum of theta0
		mbp.copy_initial_to_propse();
		reset_guess();
		mbp.propose->set_param(theta[1]);
		for(auto sett = 0u; sett < T; sett++) for(auto c = 0u; c < data.narea; c++) mbp.propose->set_transmean(sett,c);
		pk.scalar_d("mbp.c.Li",mbp.propose->likelihood());
		pk.f64("mbp.c.transmean",flat4(mbp.propose->transmean,T));
		details.stochastic = stoch0;
	}

	// ---- Monte-Carlo replicates with the reference's OWN samplers (std::poisson_distribution etc.,
	// utils.cc:148-153) for the distributional tests: totals per transition and final populations
	if(getenv("BEEP_HARNESS_REPS") != NULL && theta.size() >= 1){
		auto R = (unsigned int)atoi(getenv("BEEP_HARNESS_REPS"));
		details.stochastic = true;
		State st(details,data,model,obsmodel);
		auto J = model.trans.size(), C = model.comp.size();
		vector <double> ttot, pfin, ovs;
		for(auto r = 0u; r < R; r++){
			reset_guess();
			st.simulate(theta[0]);
			for(auto tr = 0u; tr < J; tr++){
				auto sum = 0.0;
				for(auto sett = 0u; sett < T; sett++) for(auto c = 0u; c < data.narea; c++) for(auto v : st.transnum[sett][c][tr]) sum += v;
				ttot.push_back(sum);
			}
			for(auto co = 0u; co < C; co++){
				auto sum = 0.0;
				for(auto c = 0u; c < data.narea; c++) for(auto v : st.pop[T-1][c][co]) sum += v;
				pfin.push_back(sum);
			}
			ovs.push_back(st.EF);
		}
		pk.f64("stat.trans_total",ttot,{(long)R,(long)J});
		pk.f64("stat.pop_final",pfin,{(long)R,(long)C});
		pk.f64("stat.EF",ovs);
		details.stochastic = stoch0;
	}

	pk.close();
	cout << "ref_harness: wrote " << outfile << " (" << pk.n << " arrays)" << endl;
	return 0;
}
