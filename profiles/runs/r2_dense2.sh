timeout -s KILL 400 python -m pytest tests/test_dense_gpu.py -x -q 2>&1 | tail -3
for A in 128 512 1024; do P=$((65536/A)); for dm in 0 1; do
BEEP_DENSE_M=$dm BEEP_BENCH_NNEIGH=$((A-1)) timeout -s KILL 300 python bench.py --no-cpu --no-extras --areas $A --particles $P --steps 3 --warmup 3 > gpurun_out/frontier_dense_A${A}_d$dm.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/frontier_dense_A${A}_d$dm.json'));print('A=$A P=$P dense=$dm', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', d['gpu_launches'])" || tail -3 gpurun_out/tmp.err; done; done
BEEP_DENSE_M=1 BEEP_BENCH_NNEIGH=1023 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_imap_dmma -s 20 -c 1 -o gpurun_out/r2_prof_imap_dmma -f python bench.py --no-cpu --no-extras --areas 1024 --particles 64 --steps 1 --warmup 0 > gpurun_out/r2_prof_imap.log 2>&1; tail -1 gpurun_out/r2_prof_imap.log | cut -c1-100
