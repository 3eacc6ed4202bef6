BEEP_CHUNK_NT=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traj_chunk -c 1 -o gpurun_out/r2_prof_chunk_v2_nt4 -f python bench.py --no-cpu --steps 1 --warmup 0 > gpurun_out/r2_prof_v2.log 2>&1
tail -2 gpurun_out/r2_prof_v2.log
