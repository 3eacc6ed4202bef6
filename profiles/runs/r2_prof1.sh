set -x
for nt in 4 8; do
BEEP_CHUNK_NT=$nt timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traj_chunk -c 1 -o gpurun_out/r2_prof_chunk_nt$nt -f python bench.py --no-cpu --steps 1 --warmup 0 > gpurun_out/r2_prof_nt$nt.log 2>&1
tail -3 gpurun_out/r2_prof_nt$nt.log
done
