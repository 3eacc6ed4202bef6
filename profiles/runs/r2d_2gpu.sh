#!/bin/bash
# round 2, session 3: two GPUs -- the library's NCCL collectives (tests) and the bench line as the driver launches it
cd "$GRAFT_REPO_ROOT"
timeout -s KILL 900 python -m pytest tests/test_comm_gpu.py tests/test_shard_gpu.py -q > gpurun_out/r2d_2gpu_tests.log 2>&1; tail -3 gpurun_out/r2d_2gpu_tests.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/r2d_bench_n2.json').read().strip().split('\n') if l.startswith('{')][-1]);print('N=2', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms'); print(json.dumps(d.get('filter'))[:600]); print(json.dumps(d.get('abcsmc'))[:400])" || tail -5 gpurun_out/r2d_bench_n2.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref_n2.json 2> gpurun_out/r2d_bench_ref_n2.err; tail -c 300 gpurun_out/r2d_bench_ref_n2.json
