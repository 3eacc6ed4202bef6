for nt in 4 3 2; do
BEEP_CHUNK_NT=$nt timeout -s KILL 300 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2c_nt$nt.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_nt$nt.json'));print('nt=$nt', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
done
BEEP_CHUNK_NT=3 timeout -s KILL 300 python -m pytest tests/test_fullsize_gpu.py -x -q 2>&1 | tail -3
