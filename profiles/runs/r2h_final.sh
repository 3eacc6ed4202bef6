#!/bin/bash
# round 2, session 4, after the last trim of the age contraction: the whole GPU suite, the default bench line, the launch list
cd "$GRAFT_REPO_ROOT"
( time timeout -s KILL 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2h_gpu_tests.log 2>&1; tail -4 gpurun_out/r2h_gpu_tests.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/r2h_bench_n1.json'));print('synthetic', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['value']/1e9,3), 'cpu', round(d['cpu_baseline']['value']/1e6,1), 'filter', d.get('filter',{}).get('seconds_per_filter'), 'abc', d.get('abcsmc',{}).get('s_per_generation'))" || tail -5 gpurun_out/r2h_bench_n1.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2h_launches.log 2>&1
