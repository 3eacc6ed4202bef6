#!/bin/bash
# round 2, session 3: the whole GPU suite, the default bench line, auxiliary timings (permute / copies)
cd "$GRAFT_REPO_ROOT"
( time timeout -s KILL 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2d_gpu_tests.log 2>&1; tail -8 gpurun_out/r2d_gpu_tests.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 600 gpurun_out/r2d_bench_n1.json; tail -3 gpurun_out/r2d_bench_n1.err
timeout -s KILL 600 python profiles/time_aux.py > gpurun_out/r2d_time_aux.txt 2>&1; tail -12 gpurun_out/r2d_time_aux.txt
