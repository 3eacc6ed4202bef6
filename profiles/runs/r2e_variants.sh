#!/bin/bash
# round 2, session 4: kernel variants side by side on one box (libraries under beep_b200/lib/variants, chosen by BEEP_B200_LIB)
cd "$GRAFT_REPO_ROOT"
for v in "$@"; do
  export BEEP_B200_LIB=$GRAFT_REPO_ROOT/beep_b200/lib/variants/$v.so
  echo "== $v"
  if [ "$v" != base ]; then timeout 600 python -m pytest tests/test_chunk_gpu.py tests/test_statistical_gpu.py -x -q 2>&1 | tail -1; fi
  timeout -s KILL 300 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2e_syn_$v.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2e_syn_$v.json'));print('synthetic', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
  timeout -s KILL 300 python bench.py --workload england_as --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2e_eng_$v.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2e_eng_$v.json'));print('england_as', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms')" || tail -3 gpurun_out/tmp.err
done
