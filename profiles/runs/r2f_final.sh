#!/bin/bash
# round 2, session 4: the whole GPU suite, the default bench line, the other flavours, launch list and ncu capture of the chunk kernel
cd "$GRAFT_REPO_ROOT"
( time timeout -s KILL 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2f_gpu_tests.log 2>&1; tail -4 gpurun_out/r2f_gpu_tests.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; python -c "
import json;d=json.load(open('gpurun_out/r2f_bench_n1.json'));print('synthetic', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['value']/1e9,3), 'cpu', round(d['cpu_baseline']['value']/1e6,1), 'filter', d.get('filter',{}).get('seconds_per_filter'), 'abc', d.get('abcsmc',{}).get('s_per_generation'))" || tail -5 gpurun_out/r2f_bench_n1.err
timeout -s KILL 300 python bench.py --workload england_as --no-extras --steps 5 --warmup 3 > gpurun_out/r2f_bench_england.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2f_bench_england.json'));print('england', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4), d['cpu_baseline']['kind'])" || tail -3 gpurun_out/tmp.err
timeout -s KILL 300 python bench.py --mode mbp --particles 4 --no-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2f_bench_mbp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2f_bench_mbp.json'));print('mbp', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2f_launches.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_traj_chunk -c 1 -o gpurun_out/r2f_prof_syn -f python bench.py --no-cpu --no-extras --steps 1 --warmup 0 > gpurun_out/r2f_prof_syn.log 2>&1; tail -1 gpurun_out/r2f_prof_syn.log | cut -c1-100
