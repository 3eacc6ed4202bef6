set -x
timeout 600 python -m pytest tests/test_chunk_gpu.py -x -q 2>&1 | tail -15
for nt in 8 4; do BEEP_CHUNK_NT=$nt timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r2_bench_nt$nt.json 2> gpurun_out/r2_bench_nt$nt.err; tail -c 600 gpurun_out/r2_bench_nt$nt.json; tail -3 gpurun_out/r2_bench_nt$nt.err; done
BEEP_NO_CHUNK=1 timeout 300 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r2_bench_old.json 2> gpurun_out/r2_bench_old.err; tail -c 400 gpurun_out/r2_bench_old.json
