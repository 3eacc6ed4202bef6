// ref_decisions.cc -- TEST INFRASTRUCTURE (never shipped, never on the product path).
//
// Golden vectors for the resampling DECISIONS of the particle methods, produced by
// the reference's OWN member functions on the unmodified reference objects (oracle/ref/build_ref.sh):
//     PMCMC::bootstrap      src/pmcmc.cc:211-268     -> beep_bootstrap
//     ABCMBP::EF_cutoff     src/abcmbp.cc:135-217    -> beep_ef_cutoff
//     PAS::bootstrap        src/pas.cc:127-209       -> beep_pas_bootstrap
//     MC3::swap_states      src/mc3.cc:153-217       -> beep_mc3_swap
// The reference decides with its global ran() (utils.cc:50); here ran() is replaced AT LINK TIME (utils.cc compiled
// once more with -Dran=beepref_ran, sources untouched) by the counter-based stream the library keys its decisions
// with (include/beep_rng.h: seed, generation / section / sweep, tag), so that the library must reproduce the
// reference's decision -- copy lists, ancestors, swap orders, acceptance counts -- exactly.
//
// usage: ref_decisions which out.bpk seed inputfile=... mode=... [key=value ...]     which = pmcmc | abcmbp | pas | mc3
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <iostream>
#include <random>
#include <algorithm>

#define private public
#define protected public
#include "inputs.hh"
#include "details.hh"
#include "data.hh"
#include "model.hh"
#include "obsmodel.hh"
#include "output.hh"
#include "mpi.hh"
#include "pmcmc.hh"
#include "abcmbp.hh"
#include "pas.hh"
#include "mc3.hh"
#include "utils.hh"
#include "timers.hh"
#undef private
#undef protected

#include "beep_rng.h"

using namespace std;

double beepref_ran();                                                // the original, utils.cc:50
static bool g_inject = false;
static BeepStream g_stream;
static long g_nran = 0;
double ran()
{
	if(!g_inject) return beepref_ran();
	g_nran++;
	return beep_next_uniform(&g_stream);
}
static void inject_on(uint64_t seed, uint32_t key, uint32_t tag){ g_inject = true; g_nran = 0; beep_stream_init(&g_stream,seed,key,tag,0xFFFFFFFEu); }
static void inject_off(){ g_inject = false; }

struct Pack {
	FILE *f; unsigned int n;
	Pack(const char *file){ f = fopen(file,"wb"); if(!f){ perror(file); exit(1);} n = 0; fwrite("BPK1",1,4,f); fwrite(&n,4,1,f); }
	void head(const string &name, unsigned int dtype, const vector <long> &dims){
		unsigned int len = name.size(); fwrite(&len,4,1,f); fwrite(name.data(),1,len,f);
		fwrite(&dtype,4,1,f); unsigned int nd = dims.size(); fwrite(&nd,4,1,f);
		for(auto d : dims){ long long dd = d; fwrite(&dd,8,1,f);}
		n++;
	}
	void i32(const string &name, const vector <int> &v){ head(name,0,{(long)v.size()}); if(v.size()) fwrite(v.data(),4,v.size(),f); }
	void f64(const string &name, const vector <double> &v){ head(name,1,{(long)v.size()}); if(v.size()) fwrite(v.data(),8,v.size(),f); }
	void close(){ fseek(f,4,SEEK_SET); fwrite(&n,4,1,f); fclose(f); }
};
static vector <int> toint(const vector <unsigned int> &v){ vector <int> o; for(auto x : v) o.push_back(x == UNSET ? -1 : (int)x); return o; }

int main(int argc, char **argv)
{
	if(argc < 6){ cerr << "usage: ref_decisions which out.bpk seed inputfile=... mode=... [key=value ...]" << endl; return 1;}
	string which = argv[1], outfile = argv[2];
	const uint64_t seed = (uint64_t)atoll(argv[3]);
	timersinit();
	vector <char*> av; av.push_back(argv[0]); for(int i = 4; i < argc; i++) av.push_back(argv[i]);
	Inputs inputs((int)av.size(),av.data());
	Details details(inputs);
	Mpi mpi(details);
	Data data(inputs,details,mpi);
	Model model(inputs,details,data,mpi);
	sran(7);
	ObservationModel obsmodel(details,data,model);
	Output output(details,data,model,inputs,obsmodel,mpi);
	mt19937 gen(seed);
	normal_distribution <double> nd(0.0,1.0);
	uniform_real_distribution <double> ud(0.0,1.0);
	Pack pk(outfile.c_str());
	const int ncase = 6;
	pk.i32("ncase",{ncase}); pk.f64("seed",{(double)seed});

	if(which == "pmcmc"){
		PMCMC pm(details,data,model,inputs,output,obsmodel,mpi);
		const unsigned int Ntot = pm.Ntot;
		pm.backpart.resize(ncase+2); for(auto &b : pm.backpart) b.resize(Ntot);
		for(int k = 0; k < ncase; k++){
			vector <double> L(Ntot);
			const double spread = (k%3 == 0) ? 0.5 : (k%3 == 1 ? 5.0 : 60.0);    // flat / peaked / degenerate weights
			for(auto &v : L) v = -500+spread*nd(gen);
			pm.invT = (k < 3) ? 1.0 : 0.4;
			inject_on(seed,(uint32_t)k,0xB0075u);
			auto ob = pm.bootstrap(k,L);
			inject_off();
			string p = "case"+to_string(k)+".";
			pk.f64(p+"L",L); pk.f64(p+"invT",{pm.invT}); pk.f64(p+"obprob",{ob}); pk.i32(p+"backpart",toint(pm.backpart[k+1])); pk.i32(p+"nran",{(int)g_nran});
		}
	}
	else if(which == "abcmbp"){
		ABCMBP ab(details,data,model,inputs,output,obsmodel,mpi);
		const unsigned int Ntot = ab.Ntot;
		for(int k = 0; k < ncase; k++){
			vector <Particle> part(Ntot);
			vector <double> EF(Ntot);
			for(auto i = 0u; i < Ntot; i++){ EF[i] = 1000*exp(0.7*nd(gen)); part[i].EF = EF[i]; part[i].run = 0;}
			if(k == 5){ EF[3] = EF[2]; part[3].EF = EF[2];}                       // a tie
			vector <unsigned int> pc(Ntot,12345u);
			Generation g;
			ab.EF_cutoff(g,part,pc);
			string p = "case"+to_string(k)+".";
			pk.f64(p+"EF",EF); pk.f64(p+"out",{g.EFcut,g.EFmin,g.EFmax}); pk.i32(p+"partcopy",toint(pc));
			vector <double> after; for(const auto &pa : part) after.push_back(pa.EF);
			pk.f64(p+"EF_after",after);                                           // Mpi::copy_particles applied (single rank)
		}
	}
	else if(which == "pas"){
		PAS pas(details,data,model,inputs,output,obsmodel,mpi);
		const unsigned int Ntot = pas.Ntot;
		pk.i32("nrun",{(int)pas.nrun}); pk.i32("npart",{(int)pas.npart});
		for(int k = 0; k < ncase; k++){
			vector <Particle> part(Ntot);
			vector <double> EF(Ntot);
			for(auto i = 0u; i < Ntot; i++){ EF[i] = 800*exp(0.5*nd(gen))+(i/pas.npart)*50; part[i].EF = EF[i]; part[i].run = i/pas.npart;}
			vector <unsigned int> pc(Ntot,12345u);
			Generation g;
			const double invT = 0.01*k;
			pas.quench_factor = 0.5+0.5*k;
			pas.invT_final = (k == 4) ? invT+1e-6 : UNSET;                        // the limiter of pas.cc:156-158
			inject_on(seed,(uint32_t)k,0x9A5u);
			pas.bootstrap(g,part,pc,invT);
			inject_off();
			string p = "case"+to_string(k)+".";
			pk.f64(p+"EF",EF); pk.f64(p+"in",{invT,pas.quench_factor,pas.invT_final == UNSET ? -1.0 : pas.invT_final});
			pk.f64(p+"out",{g.invT-invT,g.EFmin,g.EFmax}); pk.i32(p+"partcopy",toint(pc)); pk.i32(p+"nran",{(int)g_nran});
		}
	}
	else if(which == "mc3"){
		MC3 mc(details,data,model,inputs,output,obsmodel,mpi);
		const unsigned int Ntot = mc.Ntot, nchain = mc.nchain;
		pk.i32("nchain",{(int)nchain});
		for(int k = 0; k < ncase; k++){
			mc.part.clear(); mc.chain.clear();
			vector <double> EF(Ntot), invT(Ntot); vector <int> run(Ntot);
			for(auto i = 0u; i < Ntot; i++){
				Particle pa; pa.run = i/nchain; pa.EF = 300+40*nd(gen)+10*(i%nchain);
				mc.part.push_back(pa);
				Chain ch("c",i%nchain); ch.invT = pow(1.0-double(i%nchain)/nchain,3.0)*(0.2+0.1*k); ch.ntr = 0; ch.nac = 0;
				mc.chain.push_back(ch);
				EF[i] = pa.EF; invT[i] = ch.invT; run[i] = pa.run;
			}
			inject_on(seed,(uint32_t)k,0x3C3u);
			mc.swap_states();
			inject_off();
			// Mpi::copy_particles(part,order_tot) has been applied: the particle now at position i carries its tag
			vector <int> order(Ntot,-1), ntr, nac;
			for(auto i = 0u; i < Ntot; i++){ for(auto j = 0u; j < Ntot; j++) if(mc.part[i].EF == EF[j]) order[i] = j; ntr.push_back(mc.chain[i].ntr); nac.push_back(mc.chain[i].nac);}
			string p = "case"+to_string(k)+".";
			pk.f64(p+"EF",EF); pk.f64(p+"invT",invT); pk.i32(p+"run",run); pk.i32(p+"order",order); pk.i32(p+"ntr",ntr); pk.i32(p+"nac",nac); pk.i32(p+"nran",{(int)g_nran});
		}
	}
	else{ cerr << "unknown decision " << which << endl; return 1;}
	pk.close();
	cout << "ref_decisions: wrote " << outfile << endl;
	return 0;
}
