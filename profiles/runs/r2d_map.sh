#!/bin/bash
# round 2, session 3: mode=map through the adapter against the reference's own run
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_adapter_gpu.py -x -q -s -k "map" > gpurun_out/map_test.log 2>&1
tail -25 gpurun_out/map_test.log
