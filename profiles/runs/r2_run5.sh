for v in var/v_minb5.so var/v_minb6.so; do
for nt in 4; do
export BEEP_B200_LIB=$PWD/beep_b200/lib/$v
BEEP_CHUNK_NT=$nt timeout 300 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json;d=json.load(open('gpurun_out/tmp.json'));print('$v NT$nt', round(d['value']/1e9,3), round(d['roofline']['kernel_ms'],1), round(d['roofline']['frac'],4))"; done; done
