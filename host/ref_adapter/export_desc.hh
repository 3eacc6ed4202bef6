// export_desc.hh -- the REFERENCE-SIDE ADAPTER of the B200 engine: flattens the reference's Details + Data + Model
// (after construction, main.cc:125-151) into the BeepModelDesc arrays of include/beep_model_fields.def.
//
// This is the code a BEEP maintainer adds next to main.cc: it reads the reference's own objects
// and nothing else.  It must be included AFTER the reference's headers, with `#define private public` around them
// (Model::inf_state and friends are private, model.hh:109-110).  Two sinks are provided by the callers:
//   * a file writer (the "BPK1" container, oracle/ref/ref_harness.cc -- golden fixtures, bench inputs);
//   * DescMemory below (host/ref_adapter/beep_b200_main.cc -- the reference's drivers running on the engine).
// A sink has  scalar_i(name, int)  scalar_d(name, double)  i32(name, vector<int>, dims = {})  f64(name, vector<double>, dims = {}).
#ifndef BEEP_EXPORT_DESC_HH
#define BEEP_EXPORT_DESC_HH

#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "beep_b200.h"

static inline int xd_u2i(unsigned int v){ return v == UNSET ? -1 : (int)v; }
static inline std::vector <double> xd_flat2(const std::vector < std::vector <double> > &x)
{
	std::vector <double> v;
	for(const auto &a : x) for(auto b : a) v.push_back(b);
	return v;
}

template <class Sink>
static void beep_export_desc(Sink &pk, const Details &details, const Data &data, const Model &model)
{
	using namespace std;
	const string P = "desc.";
	auto T = details.ndivision, A = data.narea, D = data.ndemocatpos, G = data.nage, S = data.nstrain;
	unsigned int C = model.comp.size(), J = model.trans.size(), Dq = data.ndemocatpos_per_strain;

	pk.scalar_i(P+"T",T); pk.scalar_i(P+"period",details.period); pk.scalar_i(P+"spu",details.division_per_time);
	pk.scalar_i(P+"A",A); pk.scalar_i(P+"C",C); pk.scalar_i(P+"J",J); pk.scalar_i(P+"D",D); pk.scalar_i(P+"G",G);
	pk.scalar_i(P+"S",S); pk.scalar_i(P+"Dq",Dq); pk.scalar_i(P+"nparam",model.param.size());
	pk.scalar_i(P+"nspline",model.spline.size()); pk.scalar_i(P+"nobs",data.nobs); pk.scalar_i(P+"ntable",data.datatable.size());
	pk.scalar_i(P+"stochastic",details.stochastic ? 1 : 0);
	pk.scalar_i(P+"siminf_simulate",details.siminf == SIMULATE ? 1 : 0);
	pk.scalar_i(P+"infection_trans",model.infection_trans);
	pk.scalar_i(P+"geo_spline_ref",model.geo_spline_ref);
	pk.scalar_i(P+"pmcmc_obs_period",details.pmcmc_obs_period);
	pk.scalar_i(P+"n_unsupported",0);
	{ // model modifications of mode=prediction (model.cc:1440-1585); written only where they differ from "no modification"
		vector <double> tm, bm, em, sm, ss;
		if(details.mode == PREDICTION){
			const auto &mm = model.modelmod;
			bool any_t = false, any_b = false, any_e = false, any_s = false;
			for(auto sett = 0u; sett < T; sett++) for(auto c = 0u; c < A; c++){
				for(auto tr = 0u; tr < J; tr++) for(auto dp = 0u; dp < D; dp++){ auto v = mm.transmean_mult[sett][c][tr][dp]; tm.push_back(v); if(v != 1) any_t = true; }
				for(auto st = 0u; st < S; st++){ auto v = mm.beta_mult[sett][c][st]; bm.push_back(v); if(v != 1) any_b = true; v = mm.efoi_mult[sett][c][st]; em.push_back(v); if(v != 1) any_e = true; }
			}
			for(auto sp = 0u; sp < model.spline.size(); sp++) for(auto sett = 0u; sett < T; sett++){
				sm.push_back(mm.spline_mult[sp][sett]); ss.push_back(mm.spline_set[sp][sett]);
				if(mm.spline_mult[sp][sett] != 1 || mm.spline_set[sp][sett] != UNSET) any_s = true;
			}
			if(!any_t) tm.clear(); if(!any_b) bm.clear(); if(!any_e) em.clear(); if(!any_s){ sm.clear(); ss.clear(); }
		}
		pk.f64(P+"mod_transmean",tm); pk.f64(P+"mod_beta",bm); pk.f64(P+"mod_efoi",em); pk.f64(P+"mod_spline_mult",sm); pk.f64(P+"mod_spline_set",ss);
	}
	pk.scalar_d(P+"popsize",data.popsize);

	{ // demographic change, state.cc:339-376
		vector <int> info, ar, grp; vector <double> fr;
		for(const auto &dcc : data.democat_change){
			const int ncat = dcc.frac.size() ? dcc.frac[0].size() : 0;
			info.push_back(ncat); info.push_back(ar.size()); info.push_back(dcc.area.size());
			info.push_back(dcc.dp_group.size()); info.push_back(grp.size()); info.push_back(fr.size());
			for(auto c : dcc.area) ar.push_back(c);
			for(const auto &g : dcc.dp_group) for(auto dp : g) grp.push_back(dp);
			for(const auto &f : dcc.frac) for(auto v : f) fr.push_back(v);
		}
		pk.i32(P+"dcc_info",info); pk.i32(P+"dcc_area",ar); pk.i32(P+"dcc_group",grp); pk.f64(P+"dcc_frac",fr);
	}

	{ // state graphs, obsmodel.cc:113-170
		vector <int> ty, tb, fs, aoff, ar, doff, dps, poff, pobs; vector <double> fac;
		aoff.push_back(0); doff.push_back(0); poff.push_back(0);
		for(const auto &gr : data.graph){
			ty.push_back((int)gr.type); tb.push_back(gr.datatable); fac.push_back(gr.factor); fs.push_back(xd_u2i(gr.factor_spline));
			for(auto c : gr.area) ar.push_back(c);
			aoff.push_back(ar.size());
			for(auto dp : gr.dp_sel) dps.push_back(dp);
			doff.push_back(dps.size());
			for(const auto &pt : gr.point) pobs.push_back(pt.obs);
			poff.push_back(pobs.size());
		}
		pk.i32(P+"graph_step",vector <int>(1,(int)details.graph_step));
		pk.i32(P+"graph_type",ty); pk.i32(P+"graph_table",tb); pk.f64(P+"graph_factor",fac); pk.i32(P+"graph_factor_spline",fs);
		pk.i32(P+"graph_area_off",aoff); pk.i32(P+"graph_area",ar); pk.i32(P+"graph_dp_off",doff); pk.i32(P+"graph_dp",dps);
		pk.i32(P+"graph_point_off",poff); pk.i32(P+"graph_point_obs",pobs);
	}

	{ // Dirichlet groups, model.cc:2685-2695
		vector <int> off, th; vector <double> frac; off.push_back(0);
		for(const auto &dir : model.dirichlet){
			for(const auto &par : dir.param){ th.push_back(par.th); frac.push_back(par.frac);}
			off.push_back(th.size());
		}
		pk.i32(P+"dir_off",off); pk.i32(P+"dir_th",th); pk.f64(P+"dir_frac",frac);
	}

	{ // transitions, model.cc:1116-1140
		vector <int> from, to, inf, meanp, probp; vector <double> shape;
		for(auto tr = 0u; tr < J; tr++){
			const auto &t = model.trans[tr];
			from.push_back(t.from); to.push_back(t.to); inf.push_back(t.inf == TRANS_INFECTION ? 1 : 0);
			const auto &co = model.comp[t.from];
			shape.push_back(co.shape == UNSET ? 1 : co.shape);
			for(auto dp = 0u; dp < D; dp++){
				if(t.inf == TRANS_INFECTION){ meanp.push_back(-1); probp.push_back(-1);}
				else{
					meanp.push_back(co.param_mean[dp]);
					if(co.trans.size() > 1) probp.push_back(t.param_prob[dp]); else probp.push_back(-1);
				}
			}
		}
		pk.i32(P+"trans_from",from); pk.i32(P+"trans_to",to); pk.i32(P+"trans_inf",inf); pk.f64(P+"trans_shape",shape);
		pk.i32(P+"trans_mean_param",meanp,{(long)J,(long)D}); pk.i32(P+"trans_prob_param",probp,{(long)J,(long)D});
		vector <int> cinf; for(const auto &co : model.comp) cinf.push_back(co.infectivity_param);
		pk.i32(P+"comp_inf_param",cinf);
		vector <int> is; for(auto c : model.inf_state) is.push_back(c);
		pk.i32(P+"inf_state",is);
	}

	{ // splines
		vector <int> fac, off, par; vector <double> t, mf; off.push_back(0);
		for(const auto &sp : model.spline){
			fac.push_back(sp.param_factor);
			for(const auto &p : sp.p){ t.push_back(p.t); par.push_back(p.param); mf.push_back(p.multfac);}
			off.push_back(t.size());
		}
		pk.i32(P+"spline_factor_param",fac); pk.i32(P+"spline_off",off); pk.f64(P+"spline_t",t);
		pk.i32(P+"spline_param",par); pk.f64(P+"spline_multfac",mf);
	}

	{ // susceptibility, model.cc:1144-1164
		vector <int> off, par; off.push_back(0);
		for(auto dp = 0u; dp < D; dp++){
			for(auto c = 0u; c < data.ndemocat; c++){
				if(data.democat[c].sus_vari == true) par.push_back(data.democat[c].sus_param[data.democatpos[dp][c]]);
			}
			off.push_back(par.size());
		}
		pk.i32(P+"sus_off",off); pk.i32(P+"sus_param",par);
	}

	{ // area factor, model.cc:1182-1250
		vector <int> re, ae, cp, lp, lm; vector <double> cv, tp;
		if(model.region_effect.on == true){ for(auto c = 0u; c < A; c++) re.push_back(model.region_effect.area_param[c]);}
		if(data.area_effect.on == true){ for(auto c = 0u; c < A; c++) ae.push_back(model.area_effect_param_list[c]);}
		for(auto j = 0u; j < data.ncovar; j++){
			cp.push_back(model.covariate_param[j]);
			for(auto c = 0u; c < A; c++) for(auto t = 0u; t < details.period; t++) cv.push_back(data.covar[j].value[c][t]);
		}
		if(data.level_effect.on == true){
			for(auto th : model.level_effect_param_list) lp.push_back(th);
			for(auto t = 0u; t < details.period; t++) for(auto c = 0u; c < A; c++) lm.push_back(data.level_effect.param_map[t][c]);
		}
		for(auto c = 0u; c < A; c++) tp.push_back(data.area[c].total_pop);
		pk.i32(P+"region_effect_param",re); pk.i32(P+"area_effect_param",ae); pk.i32(P+"covar_param",cp);
		pk.f64(P+"covar_value",cv); pk.i32(P+"level_param",lp); pk.i32(P+"level_map",lm); pk.f64(P+"area_total_pop",tp);
	}

	{ // age mixing
		pk.f64(P+"N0",xd_flat2(data.genQ.N[0].ele),{(long)G,(long)G});
		vector <int> ms, off, ag; off.push_back(0);
		for(const auto &mm : data.genQ.matmod){
			ms.push_back(mm.spline_ref);
			for(auto a : mm.ages) ag.push_back(a);
			off.push_back(ag.size());
		}
		pk.i32(P+"matmod_spline",ms); pk.i32(P+"matmod_age_off",off); pk.i32(P+"matmod_age",ag);
		vector <int> da; for(auto dp = 0u; dp < Dq; dp++) da.push_back(data.democatpos[dp][0]);
		pk.i32(P+"dp_age",da);
	}

	{ // beta from R
		vector <int> sr, off, ar, rf; vector <double> dist; off.push_back(0);
		for(const auto &info : model.Rspline_info){
			sr.push_back(info.spline_ref);
			for(auto c : info.area) ar.push_back(c);
			off.push_back(ar.size());
			if(info.democatpos_dist.size() != Dq){ cerr << "R_dist size" << endl; exit(1);}
			for(auto v : info.democatpos_dist) dist.push_back(v);
		}
		for(auto st = 0u; st < S; st++) rf.push_back(data.strain[st].Rfactor_param);
		pk.i32(P+"R_spline_ref",sr); pk.i32(P+"R_area_off",off); pk.i32(P+"R_area",ar);
		pk.f64(P+"R_dist",dist,{(long)model.Rspline_info.size(),(long)Dq}); pk.i32(P+"strain_Rfactor_param",rf);
	}

	{ // efoi
		vector <int> sr, ref; vector <double> ad;
		for(const auto &info : model.efoispline_info){
			sr.push_back(info.spline_ref);
			if(info.efoi_agedist.size() != G){ cerr << "efoi_agedist size" << endl; exit(1);}
			for(auto v : info.efoi_agedist) ad.push_back(v);
		}
		for(auto st = 0u; st < S; st++) for(auto c = 0u; c < A; c++) ref.push_back(model.efoi_spl_ref[st][c]);
		pk.i32(P+"efoi_spline_ref",sr); pk.f64(P+"efoi_agedist",ad,{(long)model.efoispline_info.size(),(long)G});
		pk.i32(P+"efoi_ref",ref,{(long)S,(long)A});
	}

	{ // geographic mixing
		vector <int> off, to; vector <double> val, diag, fac; off.push_back(0);
		const auto &M = data.genQ.M;
		for(auto c = 0u; c < A; c++){
			for(auto j = 0u; j < M.to[c].size(); j++){ to.push_back(M.to[c][j]); val.push_back(M.val[c][j]);}
			off.push_back(to.size());
			diag.push_back(M.diag[c]); fac.push_back(data.genQ.factor[c]);
		}
		pk.i32(P+"M_off",off); pk.i32(P+"M_to",to); pk.f64(P+"M_val",val); pk.f64(P+"M_diag",diag); pk.f64(P+"M_factor",fac);
	}

	{ // initial population
		vector <double> pi;
		for(auto c = 0u; c < A; c++) for(auto co = 0u; co < C; co++) for(auto dp = 0u; dp < D; dp++) pi.push_back(data.area[c].pop_init[co][dp]);
		pk.f64(P+"pop_init",pi,{(long)A,(long)C,(long)D});
	}

	{ // observations
		vector <double> val, sd, fac, shape, thr; vector <int> fs, tab, om, si, sf, aoff, ar, doff, dp, ttoff, tt, tcoff, tc, tty;
		aoff.push_back(0); doff.push_back(0); ttoff.push_back(0); tcoff.push_back(0);
		for(const auto &ob : data.obs){
			val.push_back(ob.value); sd.push_back(ob.sd); fac.push_back(ob.factor);
			shape.push_back(ob.obsmodel == NEGBINO_OBSMODEL ? ob.shape : UNSET);   // the reference leaves shape uninitialised for the other observation models
			fs.push_back(xd_u2i(ob.factor_spline)); tab.push_back(ob.datatable); om.push_back((int)ob.obsmodel);
			si.push_back(ob.sett_i); sf.push_back(ob.sett_f);
			for(auto c : ob.area) ar.push_back(c);
			aoff.push_back(ar.size());
			for(auto d : ob.dp_sel) dp.push_back(d);
			doff.push_back(dp.size());
		}
		for(const auto &dt : data.datatable){
			for(auto tr : dt.translist) tt.push_back(tr);
			ttoff.push_back(tt.size());
			for(auto co : dt.complist) tc.push_back(co);
			tcoff.push_back(tc.size());
			thr.push_back(dt.threshold);
			tty.push_back((int)dt.type);
		}
		pk.f64(P+"obs_value",val); pk.f64(P+"obs_sd",sd); pk.f64(P+"obs_factor",fac); pk.f64(P+"obs_shape",shape);
		pk.i32(P+"obs_factor_spline",fs); pk.i32(P+"obs_table",tab); pk.i32(P+"obs_model",om);
		pk.i32(P+"obs_sett_i",si); pk.i32(P+"obs_sett_f",sf);
		pk.i32(P+"obs_area_off",aoff); pk.i32(P+"obs_area",ar); pk.i32(P+"obs_dp_off",doff); pk.i32(P+"obs_dp",dp);
		pk.i32(P+"table_trans_off",ttoff); pk.i32(P+"table_trans",tt); pk.i32(P+"table_comp_off",tcoff); pk.i32(P+"table_comp",tc);
		pk.f64(P+"table_threshold",thr); pk.i32(P+"table_type",tty);
	}
}


/// In-memory sink: owns the arrays and hands out the BeepModelDesc that points into them.
struct DescMemory {
	std::map <std::string, std::vector <int> > I;
	std::map <std::string, std::vector <double> > F;
	static std::string strip(const std::string &n){ return n.compare(0,5,"desc.") == 0 ? n.substr(5) : n; }
	void scalar_i(const std::string &name, int v){ I[strip(name)] = std::vector <int>(1,v); }
	void scalar_d(const std::string &name, double v){ F[strip(name)] = std::vector <double>(1,v); }
	void i32(const std::string &name, const std::vector <int> &v, std::vector <long> = {}){ I[strip(name)] = v; }
	void f64(const std::string &name, const std::vector <double> &v, std::vector <long> = {}){ F[strip(name)] = v; }
	BeepModelDesc desc()
	{
		BeepModelDesc d;
		memset(&d,0,sizeof d);
#define BEEP_I32(name) d.name = I.at(#name)[0];
#define BEEP_F64(name) d.name = F.at(#name)[0];
#define BEEP_AI32(name) { auto &a = I[#name]; d.name = a.data(); d.n_##name = (int64_t)a.size(); }
#define BEEP_AF64(name) { auto &a = F[#name]; d.name = a.data(); d.n_##name = (int64_t)a.size(); }
#include "beep_model_fields.def"
#undef BEEP_I32
#undef BEEP_F64
#undef BEEP_AI32
#undef BEEP_AF64
		return d;
	}
};

#endif
