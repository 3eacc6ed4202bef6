#!/bin/bash
# round 2, session 3: ncu capture of k_likelihood_approx (EX_C5: 128 compartments, one CTA per SM)
cd "$GRAFT_REPO_ROOT"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_likelihood_approx -c 1 -o gpurun_out/r2d_approx_c5 -f python profiles/prof_approx.py APX_C5 148 > gpurun_out/r2d_approx_c5.log 2>&1; tail -2 gpurun_out/r2d_approx_c5.log | cut -c1-200
