#!/bin/bash
# round 2, session 3: final ncu capture of k_likelihood_approx, timings of the auxiliary kernels, graph tests
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_graphs_gpu.py tests/test_decisions_cpu.py -x -q 2>&1 | tail -2
timeout -s KILL 600 python profiles/time_aux.py > gpurun_out/r2d_time_aux.txt 2>&1; tail -6 gpurun_out/r2d_time_aux.txt
timeout 600 python profiles/time_approx.py > gpurun_out/time_approx.log 2>&1; tail -3 gpurun_out/time_approx.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_likelihood_approx -c 1 -o gpurun_out/r2d_approx_c5 -f python profiles/prof_approx.py APX_C5 148 > gpurun_out/r2d_approx_c5.log 2>&1; tail -1 gpurun_out/r2d_approx_c5.log | cut -c1-120
