#!/bin/bash
# round 2, session 3: the whole GPU suite, the default bench line (both arms), launch list
cd "$GRAFT_REPO_ROOT"
( time timeout -s KILL 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_gpu_tests.log 2>&1; tail -5 gpurun_out/r2d_gpu_tests.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; tail -c 1500 gpurun_out/r2d_bench_n1.json; tail -3 gpurun_out/r2d_bench_n1.err
( time timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; tail -c 800 gpurun_out/r2d_bench_ref.json; tail -3 gpurun_out/r2d_bench_ref.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2d_launches.log 2>&1; tail -2 gpurun_out/r2d_launches.log | cut -c1-200
