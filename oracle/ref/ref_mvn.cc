// ref_mvn.cc -- TEST INFRASTRUCTURE (never shipped, never on the product path).
//
// Golden vectors for the ABC-SMC importance-weight kernel, from the
// UNMODIFIED reference objects (oracle/ref/build_ref.sh): the reference's own MVN::setup
// (weighted covariance, mvn.cc:96-129; Cholesky + invert_matrix, :148-181), MVN::propose and
// MVN::probability (mvn.cc:185-196) drive the denominator of ABCSMC::calculate_particle_weight
// (abcsmc.cc:260-271):   sum_i = SUM_j w_j exp(mvn.probability(prop_i, last_j)).
//
// usage: ref_mvn out.bpk nparam nvar Nlast Nprop seed        (BPK1 container, see beep_b200/bpk.py)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <random>
#include <iostream>

#define private public
#include "mvn.hh"
#undef private
#include "utils.hh"
#include "timers.hh"

using namespace std;

struct Pack {
	FILE *f; unsigned int n;
	Pack(const char *file){ f = fopen(file,"wb"); if(!f){ perror(file); exit(1);} n = 0; fwrite("BPK1",1,4,f); fwrite(&n,4,1,f); }
	void head(const string &name, unsigned int dtype, const vector <long> &dims){
		unsigned int len = name.size(); fwrite(&len,4,1,f); fwrite(name.data(),1,len,f);
		fwrite(&dtype,4,1,f); unsigned int nd = dims.size(); fwrite(&nd,4,1,f);
		for(auto d : dims){ long long dd = d; fwrite(&dd,8,1,f);}
		n++;
	}
	void i32(const string &name, const vector <int> &v, vector <long> dims){ head(name,0,dims); fwrite(v.data(),4,v.size(),f); }
	void f64(const string &name, const vector <double> &v, vector <long> dims){ head(name,1,dims); fwrite(v.data(),8,v.size(),f); }
	void close(){ fseek(f,4,SEEK_SET); fwrite(&n,4,1,f); fclose(f); }
};

int main(int argc, char **argv)
{
	if(argc < 7){ fprintf(stderr,"usage: %s out.bpk nparam nvar Nlast Nprop seed\n",argv[0]); return 2; }
	const unsigned int nparam = atoi(argv[2]), nvar = atoi(argv[3]), Nl = atoi(argv[4]), N = atoi(argv[5]);
	const int seed = atoi(argv[6]);
	timersinit();                                            // invert_matrix times itself (main.cc:109)
	sran(seed);                                              // the reference's generators (utils.cc:41-62)
	mt19937 gen(seed);
	normal_distribution <double> nd(0.0,1.0);
	uniform_real_distribution <double> ud(0.0,1.0);

	// the free parameters: every parameter whose index is not a multiple of 5 until nvar are taken
	vector <unsigned int> var;
	for(auto th = 0u; th < nparam && var.size() < nvar; th++) if(th%5 != 0) var.push_back(th);
	if(var.size() != nvar){ fprintf(stderr,"nparam too small for nvar\n"); return 2; }

	// correlated samples of the last generation: a few shared factors + own noise, different scales
	const unsigned int nfac = 4;
	vector <double> base(nparam), scale(nparam);
	vector < vector <double> > load(nparam, vector <double>(nfac));
	for(auto th = 0u; th < nparam; th++){
		base[th] = 0.2 + 5*ud(gen); scale[th] = 0.01 + 0.5*ud(gen)*base[th];
		for(auto k = 0u; k < nfac; k++) load[th][k] = 0.6*nd(gen);
	}
	vector <ParamSample> samp(Nl);
	vector <double> w(Nl);
	for(auto j = 0u; j < Nl; j++){
		vector <double> z(nfac); for(auto &v : z) v = nd(gen);
		samp[j].run = 0; samp[j].EF = 0; samp[j].paramval.resize(nparam);
		for(auto th = 0u; th < nparam; th++){
			double x = nd(gen); for(auto k = 0u; k < nfac; k++) x += load[th][k]*z[k];
			samp[j].paramval[th] = base[th] + scale[th]*x;
		}
		w[j] = 0.05 + ud(gen);
	}

	MVN mvn("ABCSMC",var,1.3,ALL_PARAM,MULTIPLE);            // abcsmc.cc:28 with propsize = 1.3
	mvn.setup(samp,w);                                       // abcsmc.cc:55

	vector < vector <double> > prop(N);
	vector <double> sums(N);
	for(auto i = 0u; i < N; i++){
		auto p = (unsigned int)(ud(gen)*Nl); if(p >= Nl) p = Nl-1;
		prop[i] = mvn.propose(samp[p].paramval);               // abcsmc.cc:68
		auto sum = 0.0;                                        // abcsmc.cc:264-268
		for(auto j = 0u; j < Nl; j++) sum += w[j]*exp(mvn.probability(prop[i],samp[j].paramval));
		sums[i] = sum;
	}

	Pack pk(argv[1]);
	vector <int> vi(var.begin(),var.end());
	pk.i32("var",vi,{(long)nvar});
	pk.f64("size",vector <double>{mvn.size},{1});
	vector <double> flat;
	for(auto &r : mvn.M) for(auto v : r) flat.push_back(v);
	pk.f64("M",flat,{(long)nvar,(long)nvar});
	flat.clear(); for(auto &r : mvn.inv_M) for(auto v : r) flat.push_back(v);
	pk.f64("inv_M",flat,{(long)nvar,(long)nvar});
	flat.clear(); for(auto &s : samp) for(auto v : s.paramval) flat.push_back(v);
	pk.f64("param_last",flat,{(long)Nl,(long)nparam});
	pk.f64("w_last",w,{(long)Nl});
	flat.clear(); for(auto &s : prop) for(auto v : s) flat.push_back(v);
	pk.f64("param_prop",flat,{(long)N,(long)nparam});
	pk.f64("sums",sums,{(long)N});
	pk.close();
	return 0;
}
