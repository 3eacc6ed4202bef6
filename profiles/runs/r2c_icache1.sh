timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout -s KILL 300 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2c_ic_syn.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_ic_syn.json'));print('synthetic', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
timeout -s KILL 300 python bench.py --workload england_as --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2c_ic_eng.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_ic_eng.json'));print('england', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4), d['roofline'].get('kernel_ms'))" || tail -3 gpurun_out/tmp.err
timeout -s KILL 300 python bench.py --mode mbp --particles 4 --no-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2c_ic_mbp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_ic_mbp.json'));print('mbp', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_traj -c 1 -o gpurun_out/r2c_prof_eng -f python bench.py --workload england_as --no-cpu --no-extras --steps 1 --warmup 0 > gpurun_out/r2c_prof_eng.log 2>&1; tail -1 gpurun_out/r2c_prof_eng.log | cut -c1-100
