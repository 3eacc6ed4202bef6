#!/bin/bash
# round 2, session 4: the final library under torchrun on two GPUs (bench line with the filter's NCCL collectives)
cd "$GRAFT_REPO_ROOT"
timeout -s KILL 85 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; python -c "
import json;d=json.load(open('gpurun_out/r2h_bench_n2.json'));print('n2', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', 'filter', d.get('filter',{}).get('seconds_per_filter'), d.get('filter',{}).get('collectives_share'), 'abc', d.get('abcsmc',{}).get('s_per_generation'))" || tail -5 gpurun_out/r2h_bench_n2.err
