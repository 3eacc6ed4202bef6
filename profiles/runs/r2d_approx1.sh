#!/bin/bash
# round 2, session 3: first GPU run of the MVN-approximate likelihood (row N3)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_approx_gpu.py -x -q > gpurun_out/approx_tests.log 2>&1
tail -15 gpurun_out/approx_tests.log
timeout 600 python profiles/time_approx.py > gpurun_out/time_approx.log 2>&1
cat gpurun_out/time_approx.log
