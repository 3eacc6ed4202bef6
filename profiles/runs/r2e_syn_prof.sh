#!/bin/bash
# round 2, session 4: ncu capture of the chunk kernel at the bench configuration
cd "$GRAFT_REPO_ROOT"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_traj_chunk -c 1 -o gpurun_out/r2e_prof_syn -f python bench.py --no-cpu --no-extras --steps 1 --warmup 0 > gpurun_out/r2e_prof_syn.log 2>&1; tail -1 gpurun_out/r2e_prof_syn.log | cut -c1-100
