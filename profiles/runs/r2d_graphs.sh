#!/bin/bash
# round 2, session 3: graph states (row N4) + approx tests on the device
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_graphs_gpu.py tests/test_approx_gpu.py -x -q > gpurun_out/graphs_tests.log 2>&1
tail -25 gpurun_out/graphs_tests.log
