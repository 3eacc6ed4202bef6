timeout -s KILL 300 python -m pytest tests/test_adapter_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -6
(time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err); tail -5 gpurun_out/r2_bench_n2.err; tail -1 gpurun_out/r2_bench_n2.json > gpurun_out/r2_bench_n2.line; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n2.line').read())
print(json.dumps({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}))
print(json.dumps(d.get('filter'),indent=0)); print(json.dumps(d.get('abcsmc'),indent=0))"
