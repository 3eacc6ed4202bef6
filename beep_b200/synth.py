"""Synthetic scale-up populations (BASELINE.json configs[4]).

Builds a flat ``ModelDesc`` of the named shape -- A areas x G age groups x 12
compartments (S, E, A, I, C, H, R, D with Erlang-2 stages on E, A, I, H) x 13
transitions -- with the same normalisations the reference applies to real data
(geographic mixing matrix: src/data_matrix.cc:55-96; age mixing matrix:
:100-134; "normal p%" observation sd: src/data.cc:1589-1623).  The inputs are
synthetic and seeded; nothing here is read from the reference at run time.
"""
import numpy as np

from .model import ModelDesc

SEED = 20261017
COMPS = ["S", "E0", "E1", "A0", "A1", "I0", "I1", "C", "H0", "H1", "R", "D"]
# (from, to); transition 0 is the infection
TRANS = [("S", "E0"), ("E0", "E1"), ("E1", "A0"), ("E1", "I0"), ("A0", "A1"), ("A1", "R"), ("I0", "I1"),
         ("I1", "R"), ("I1", "C"), ("C", "H0"), ("H0", "H1"), ("H1", "R"), ("H1", "D")]
UNSET_D = 999999999.0


def make_model(A=16, G=20, T=365, spu=1, nneigh=8, seed=SEED, obs=True, mean_pop=5e3, stochastic=1):
    """Returns (ModelDesc, theta_true[nparam])."""
    rng = np.random.default_rng(seed)
    C, J, D, S = len(COMPS), len(TRANS), G, 1
    period = T // spu
    assert period * spu == T
    ci = {n: i for i, n in enumerate(COMPS)}

    # ---- parameters -----------------------------------------------------
    names, vals = [], []

    def par(name, v):
        names.append(name)
        vals.append(float(v))
        return len(vals) - 1

    p_zero, p_infI, p_infA, p_one = par("zero", 0), par("inf_I", 1.0), par("inf_A", 0.55), par("one", 1.0)
    p_mean = {"E": par("mean_E", 3.0), "A": par("mean_A", 5.0), "I": par("mean_I", 4.0), "C": par("mean_C", 2.0), "H": par("mean_H", 8.0)}
    nknot = 12
    p_R = [par(f"R_{k}", rng.uniform(0.7, 2.5)) for k in range(nknot)]
    p_efoi = par("efoi", 10.0)
    age = np.arange(G) / max(G - 1, 1)
    p_sus = [par(f"sus_{a}", 0.7 + 0.6 * rng.uniform()) for a in range(G)]
    pA = 0.6 - 0.35 * age
    pC = 0.02 + 0.25 * age ** 2
    pD = 0.02 + 0.4 * age ** 2
    p_bA = [par(f"pA_{a}", pA[a]) for a in range(G)]
    p_bI = [par(f"pI_{a}", 1 - pA[a]) for a in range(G)]
    p_bR = [par(f"pIR_{a}", 1 - pC[a]) for a in range(G)]
    p_bC = [par(f"pIC_{a}", pC[a]) for a in range(G)]
    p_bHR = [par(f"pHR_{a}", 1 - pD[a]) for a in range(G)]
    p_bHD = [par(f"pHD_{a}", pD[a]) for a in range(G)]
    theta = np.array(vals)
    nparam = theta.size

    # ---- compartments / transitions --------------------------------------
    shape = {"E0": 2, "E1": 2, "A0": 2, "A1": 2, "I0": 2, "I1": 2, "H0": 2, "H1": 2, "C": 1}
    comp_inf = np.full(C, p_zero, dtype=np.int32)
    for n in ("I0", "I1"):
        comp_inf[ci[n]] = p_infI
    for n in ("A0", "A1"):
        comp_inf[ci[n]] = p_infA
    trans_from = np.array([ci[f] for f, _ in TRANS], dtype=np.int32)
    trans_to = np.array([ci[t] for _, t in TRANS], dtype=np.int32)
    trans_inf = np.zeros(J, dtype=np.int32)
    trans_inf[0] = 1
    trans_shape = np.ones(J)
    mean_param = np.full((J, D), -1, dtype=np.int32)
    prob_param = np.full((J, D), -1, dtype=np.int32)
    branch = {("E1", "A0"): p_bA, ("E1", "I0"): p_bI, ("I1", "R"): p_bR, ("I1", "C"): p_bC, ("H1", "R"): p_bHR, ("H1", "D"): p_bHD}
    for tr, (f, t) in enumerate(TRANS):
        if tr == 0:
            continue
        trans_shape[tr] = shape[f]
        mean_param[tr, :] = p_mean[f[0]]
        if (f, t) in branch:
            prob_param[tr, :] = branch[(f, t)]
    inf_state = np.array([ci[n] for n in ("E0", "E1", "A0", "A1", "I0", "I1")], dtype=np.int32)

    # ---- splines: 0 = R(t), 1 = efoi, 2 = geographic mixing (constant 1) ---
    knots_t = np.linspace(0, period, nknot)
    spline_off = np.array([0, nknot, nknot + 2, nknot + 4], dtype=np.int32)
    spline_t = np.concatenate([knots_t, [0, period], [0, period]])
    spline_param = np.array(p_R + [p_efoi, p_efoi, p_one, p_one], dtype=np.int32)
    spline_multfac = np.ones(spline_t.size)
    spline_factor = np.array([p_one, p_one, p_one], dtype=np.int32)

    # ---- populations -------------------------------------------------------
    area_pop = np.maximum(np.round(rng.lognormal(np.log(mean_pop), 1.0, A)), 200).astype(np.int64)
    agefrac = rng.dirichlet(5 * np.ones(G))
    pop_init = np.zeros((A, C, D))
    for c in range(A):
        n = np.floor(area_pop[c] * agefrac).astype(np.int64)
        n[0] += area_pop[c] - n.sum()
        pop_init[c, ci["S"], :] = n
    # seed infections in ~2% of the areas
    nseed = max(1, A // 50)
    for c in rng.choice(A, nseed, replace=False):
        k = np.minimum(pop_init[c, ci["S"], :], 2)
        pop_init[c, ci["S"], :] -= k
        pop_init[c, ci["I0"], :] += k
    total_pop = pop_init.sum(axis=(1, 2))
    popsize = float(total_pop.sum())

    # ---- age mixing matrix, normalised as data_matrix.cc:100-134 -----------
    N0 = np.abs(rng.normal(size=(G, G)))
    N0 = 0.5 * (N0 + N0.T) + 3 * np.eye(G)
    popage = pop_init.sum(axis=(0, 1))
    ssum = 0.0
    for a in range(G):
        ssum += float(np.dot(popage, N0[:, a]))
    ssum /= G * popage.sum()
    N0 = N0 / ssum

    # ---- sparse geographic mixing: ring + random neighbours, data_matrix.cc:55-96 ----
    M_off = [0]
    M_to, M_val = [], []
    M_diag = np.ones(A)
    for c in range(A):
        nb = set()
        if A > 1:
            nb.update([(c - 1) % A, (c + 1) % A])
            k = min(nneigh, A - 1)
            while len(nb) < k:
                nb.add(int(rng.integers(0, A)))
            nb.discard(c)
        nb = sorted(nb)
        w = rng.uniform(0.01, 0.1, len(nb)) / np.maximum(total_pop[nb], 1) * total_pop[c] if nb else np.zeros(0)
        ssum = M_diag[c] * total_pop[c] + float(np.dot(w, total_pop[nb])) if nb else M_diag[c] * total_pop[c]
        M_diag[c] /= ssum
        w = w / ssum
        M_to += nb
        M_val += list(w)
        M_off.append(len(M_to))
    M_factor = (1.0 / total_pop) / M_diag

    # ---- observations: weekly infection counts + weekly hospital population per area ----
    ov, osd, ofac, oshape, ofs, otab, omod, osi, osf = [], [], [], [], [], [], [], [], []
    oa_off, oa, od_off, od = [0], [], [0], []
    if obs:
        week = 7 * spu
        nweek = T // week
        for tab in (0, 1):
            for c in range(A):
                for w in range(nweek):
                    if tab == 0:
                        si, sf = w * week, (w + 1) * week
                        val = total_pop[c] * 0.004 * (1 + np.sin(w / 8.0) ** 2)
                    else:
                        si, sf = w * week, w * week + 1
                        val = total_pop[c] * 0.0005 * (1 + np.sin(w / 8.0) ** 2)
                    ov.append(val)
                    osi.append(si)
                    osf.append(sf)
                    ofac.append(1.0)
                    oshape.append(UNSET_D)
                    ofs.append(-1)
                    otab.append(tab)
                    omod.append(1)          # NORMAL_PERCENT_OBSMODEL
                    oa.append(c)
                    oa_off.append(len(oa))
                    od += list(range(D))
                    od_off.append(len(od))
        ov = np.array(ov)
        # data.cc:1617-1621: sd = (y (1-eps) + ymax eps) p/100, eps = 0.01, p = 10
        ymax = np.array([ov[np.array(otab) == t].max() for t in (0, 1)])
        osd = (ov * 0.99 + ymax[np.array(otab)] * 0.01) * 0.1
    nobs = len(ov)
    ntable = 2
    table_trans_off = np.array([0, 1, 1], dtype=np.int32)
    table_trans = np.array([0], dtype=np.int32)
    table_comp_off = np.array([0, 0, 2], dtype=np.int32)
    table_comp = np.array([ci["H0"], ci["H1"]], dtype=np.int32)

    i32 = lambda x: np.asarray(x, dtype=np.int32)
    f64 = lambda x: np.asarray(x, dtype=np.float64)
    arrays = dict(
        T=T, period=period, spu=spu, A=A, C=C, J=J, D=D, G=G, S=S, Dq=D, nparam=nparam, nspline=3, nobs=nobs, ntable=ntable,
        stochastic=stochastic, siminf_simulate=0, infection_trans=0, geo_spline_ref=2,
        pmcmc_obs_period=7 * spu * max(1, (T // (7 * spu)) // 10), n_unsupported=0, popsize=popsize,
        dir_off=i32([0]), dir_th=i32([]), dir_frac=f64([]),
        trans_from=trans_from, trans_to=trans_to, trans_inf=trans_inf, trans_shape=trans_shape,
        trans_mean_param=mean_param, trans_prob_param=prob_param, comp_inf_param=comp_inf, inf_state=inf_state,
        spline_factor_param=spline_factor, spline_off=spline_off, spline_t=f64(spline_t), spline_param=spline_param, spline_multfac=spline_multfac,
        sus_off=i32(np.arange(D + 1)), sus_param=i32(p_sus),
        region_effect_param=i32([]), area_effect_param=i32([]), covar_param=i32([]), covar_value=f64([]),
        level_param=i32([]), level_map=i32([]), area_total_pop=f64(total_pop),
        N0=f64(N0), matmod_spline=i32([]), matmod_age_off=i32([0]), matmod_age=i32([]), dp_age=i32(np.arange(D)),
        R_spline_ref=i32([0]), R_area_off=i32([0, A]), R_area=i32(np.arange(A)), R_dist=f64(popage / popage.sum()).reshape(1, D),
        strain_Rfactor_param=i32([p_one]),
        efoi_spline_ref=i32([1]), efoi_agedist=f64(np.full((1, G), 1.0 / G)), efoi_ref=i32(np.zeros((S, A))),
        M_off=i32(M_off), M_to=i32(M_to), M_val=f64(M_val), M_diag=f64(M_diag), M_factor=f64(M_factor),
        pop_init=f64(pop_init),
        obs_value=f64(ov), obs_sd=f64(osd), obs_factor=f64(ofac), obs_shape=f64(oshape), obs_factor_spline=i32(ofs),
        obs_table=i32(otab), obs_model=i32(omod), obs_sett_i=i32(osi), obs_sett_f=i32(osf),
        obs_area_off=i32(oa_off), obs_area=i32(oa), obs_dp_off=i32(od_off), obs_dp=i32(od),
        table_trans_off=table_trans_off, table_trans=table_trans, table_comp_off=table_comp_off, table_comp=table_comp,
        table_threshold=f64([UNSET_D, UNSET_D]),
    )
    return ModelDesc(arrays), theta


def jitter_params(theta, nparticle, rel=0.05, seed=1):
    """Per-chain parameter vectors: +-rel uniform jitter around the truth, keeping
    branching probabilities of each source normalised."""
    rng = np.random.default_rng(seed)
    th = np.repeat(theta[None, :], nparticle, axis=0)
    th *= 1 + rel * rng.uniform(-1, 1, th.shape)
    th[:, 0] = 0.0
    th[:, 3] = 1.0
    return th


def write_reference_inputs(workdir, A=256, G=20, T=365, seed=SEED, name="synthetic"):
    """Writes the synthetic population of make_model(A, G, T) as input files in the REFERENCE's own
    formats (TOML + CSV: ages / comps / trans / R_spline / efoi_spline / areas / mixing matrices / data_tables,
    as in the reference's examples/EX_C2.toml and EX_C4.toml), so that the unmodified reference can run the
    same shape: A areas x G ages x 12 compartments (Erlang-2 on E, A, I, H) x 13 transitions x T daily steps.
    Same generator and seed as make_model: populations, age matrix, sparse geographic matrix (written dense,
    zeros elsewhere), branching probabilities by age and the R(t) knots are identical; age-specific
    susceptibility and the seeded infections are left at the reference's defaults (1 and none -- the
    external force of infection starts the epidemic).  Returns the TOML path (relative names inside)."""
    import os
    m, theta = make_model(A=A, G=G, T=T, seed=seed)
    ar = m.arrays
    data = os.path.join(workdir, f"Data_{name}")
    os.makedirs(data, exist_ok=True)
    ages = [f"a{k}" for k in range(G)]
    pop = np.asarray(ar["pop_init"]).reshape(A, -1, G).sum(axis=1)
    with open(os.path.join(data, "population.csv"), "w") as f:
        f.write("area," + ",".join(ages) + "\n")
        for c in range(A):
            f.write(f"R{c}," + ",".join(str(int(v)) for v in pop[c]) + "\n")
    np.savetxt(os.path.join(data, "age_matrix.csv"), np.asarray(ar["N0"]).reshape(G, G), delimiter=",", fmt="%.10g")
    geo = np.zeros((A, A))
    off, to, val, diag = ar["M_off"], ar["M_to"], ar["M_val"], ar["M_diag"]
    for c in range(A):
        geo[c, c] = diag[c]
        for k in range(off[c], off[c + 1]):
            geo[c, to[k]] = val[k]
    geo = 0.5 * (geo + geo.T) / geo.diagonal().mean()          # the reference wants a symmetric matrix and normalises it itself
    np.savetxt(os.path.join(data, "geo_matrix.csv"), geo, delimiter=",", fmt="%.10g")
    age = np.arange(G) / max(G - 1, 1)
    pA, pC, pD = 0.6 - 0.35 * age, 0.02 + 0.25 * age ** 2, 0.02 + 0.4 * age ** 2
    fmt = lambda v: " | ".join(f"{x:.6g}" for x in v)
    nknot = 12
    Rv = theta[9:9 + nknot]
    bps = ["start"] + [str(int(round(x))) for x in np.linspace(0, T, nknot)[1:-1]] + ["end"]      # break points are whole days
    toml = f'''description = "synthetic scale-up population ({A} areas x {G} ages), written by beep_b200/synth.py"
time_format = "number"
start = "0"
end = "{T}"
steps_per_unit_time = 1
datadir = "Data_{name}"
outputdir = "Output_{name}"
ages = {{cats="{' | '.join(ages)}"}}
age_mixing_matrix = "age_matrix.csv"
geo_mixing_matrix = "geo_matrix.csv"
comps = [
{{name="S"}},
{{name="E", dist="Erlang", k="2", mean_value="3"}},
{{name="A", dist="Erlang", k="2", mean_value="5", inf_value="0.55"}},
{{name="I", dist="Erlang", k="2", mean_value="4", inf_value="1"}},
{{name="C", dist="Exp", mean_value="2"}},
{{name="H", dist="Erlang", k="2", mean_value="8"}},
{{name="R"}},
{{name="D"}}
]
trans = [
{{from="S", to="E", infection="yes"}},
{{from="E", to="A", prob_value="age: {fmt(pA)}"}},
{{from="E", to="I", prob_value="age: *"}},
{{from="A", to="R"}},
{{from="I", to="R", prob_value="age: {fmt(1 - pC)}"}},
{{from="I", to="C", prob_value="age: *"}},
{{from="C", to="H"}},
{{from="H", to="R", prob_value="age: {fmt(1 - pD)}"}},
{{from="H", to="D", prob_value="age: *"}}
]
R_spline = [{{ value="{' | '.join(f'{x:.6g}' for x in Rv)}", bp="{' | '.join(bps)}"}}]
efoi_spline = [{{ value="10"}}]
areas = "population.csv"
data_tables = [
{{type="transition", observation="S->E", timestep="7", obsmodel="normal 10%", geo_dep="area", file="S-E.csv"}},
{{type="population", observation="H", timestep="7", obsmodel="normal 10%", geo_dep="area", file="H.csv"}}
]
'''
    path = os.path.join(workdir, f"{name}.toml")
    with open(path, "w") as f:
        f.write(toml)
    return path
