#!/bin/bash
# round 2, session 4: chunk-kernel variants side by side (synthetic simulate + MBP sweep; chunk and full-size parity tests)
cd "$GRAFT_REPO_ROOT"
for v in "$@"; do
  export BEEP_B200_LIB=$GRAFT_REPO_ROOT/beep_b200/lib/variants/$v.so
  echo "== $v"
  if [ "$v" != base ]; then timeout 600 python -m pytest tests/test_chunk_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -1; fi
  timeout -s KILL 300 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2g_syn_$v.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2g_syn_$v.json'));print('synthetic', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
  timeout -s KILL 300 python bench.py --mode mbp --particles 4 --no-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2g_mbp_$v.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2g_mbp_$v.json'));print('mbp', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
done
