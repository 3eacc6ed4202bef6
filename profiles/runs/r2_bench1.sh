timeout -s KILL 300 python -m pytest tests/test_adapter_gpu.py -x -q 2>&1 | tail -8
(time timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err); tail -5 gpurun_out/r2_bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_full.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','roofline','e2e')},indent=0)[:1500])
print(json.dumps(d.get('filter'),indent=0)); print(json.dumps(d.get('abcsmc'),indent=0)); print(json.dumps(d.get('cpu_baseline'))[:600])"
