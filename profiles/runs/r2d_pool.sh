#!/bin/bash
# round 2, session 3: chunk kernel after the optimistic pooling of the draw list
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_chunk_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -2
timeout -s KILL 300 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2d_synthetic.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2d_synthetic.json'));print('synthetic', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
timeout -s KILL 300 python bench.py --mode mbp --particles 4 --no-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2d_mbp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2d_mbp.json'));print('mbp', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
