#!/bin/bash
# round 2, session 3: eight GPUs -- the bench line as the driver launches it (extras included)
cd "$GRAFT_REPO_ROOT"
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/r2d_bench_n8.json').read().strip().split('\n') if l.startswith('{')][-1]);print('N=8', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms'); print(json.dumps(d.get('filter'))[:600]); print(json.dumps(d.get('abcsmc'))[:400])" || tail -8 gpurun_out/r2d_bench_n8.err
tail -4 gpurun_out/r2d_bench_n8.err
