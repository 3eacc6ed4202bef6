#!/bin/bash
# round 2, session 3: England_AS + synthetic bench after a code-size change (no profile)
cd "$GRAFT_REPO_ROOT"
for w in england_as synthetic; do
timeout -s KILL 300 python bench.py --workload $w --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2d_$w.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2d_$w.json'));print('$w', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4), d['roofline'].get('kernel_ms'))" || tail -3 gpurun_out/tmp.err
done
timeout 600 python -m pytest tests/test_statistical_gpu.py tests/test_chunk_gpu.py -x -q 2>&1 | tail -2
