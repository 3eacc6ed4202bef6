timeout -s KILL 400 python -m pytest tests/test_dense_gpu.py tests/test_chunk_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -8
for A in 128 512 1024; do P=$((65536/A)); for dm in 0 1; do
BEEP_DENSE_M=$dm BEEP_BENCH_NNEIGH=$((A-1)) timeout -s KILL 300 python bench.py --no-cpu --no-extras --areas $A --particles $P --steps 3 --warmup 3 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/tmp.json'));print('A=$A P=$P dense=$dm', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', d['gpu_launches'])" || tail -3 gpurun_out/tmp.err; done; done
