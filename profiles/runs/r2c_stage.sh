timeout -s KILL 600 python -m pytest tests/test_chunk_gpu.py tests/test_fullsize_gpu.py tests/test_dense_gpu.py -x -q 2>&1 | tail -4
for n in default g8; do
if [ $n = g8 ]; then export BEEP_B200_LIB=$PWD/beep_b200/lib/exp/lib_g8.so; fi
timeout -s KILL 300 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2c_stage_$n.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_stage_$n.json'));print('$n', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
done
unset BEEP_B200_LIB
timeout -s KILL 300 python bench.py --mode mbp --particles 4 --no-cpu --no-extras --steps 3 --warmup 3 > gpurun_out/r2c_stage_mbp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_stage_mbp.json'));print('mbp', round(d['value']/1e9,3),'G units/s', round(d['ms_per_step'],1),'ms', round(d['roofline']['frac'],4))" || tail -3 gpurun_out/tmp.err
