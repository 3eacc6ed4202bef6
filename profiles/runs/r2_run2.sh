set -x
timeout 600 python -m pytest tests/test_chunk_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -5
for nt in 4 8; do BEEP_CHUNK_NT=$nt timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r2b_bench_nt$nt.json 2> gpurun_out/r2b_bench_nt$nt.err; python -c "
import json;d=json.load(open('gpurun_out/r2b_bench_nt$nt.json'));print('NT$nt', d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
