timeout 600 python -m pytest tests/test_chunk_gpu.py -x -q 2>&1 | tail -3
for nt in 4 8; do
BEEP_CHUNK_NT=$nt timeout 300 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/tmp.json'));print('NT$nt', round(d['value']/1e9,3), round(d['roofline']['kernel_ms'],1), round(d['roofline']['frac'],4))"; done
