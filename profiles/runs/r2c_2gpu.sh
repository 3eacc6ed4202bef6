nvidia-smi -L | head -3
timeout -s KILL 900 python -m pytest tests/test_comm_gpu.py tests/test_shard_gpu.py -x -q 2>&1 | tail -5
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/r2c_bench_n2.json').read().strip().split('\n')[-1]);print('N=2', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms'); print(json.dumps(d.get('filter'))[:700]); print(json.dumps(d.get('abcsmc'))[:400])" || tail -5 gpurun_out/r2c_bench_n2.err
