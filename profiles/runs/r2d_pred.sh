#!/bin/bash
# round 2, session 3: prediction-mode model modifications on the device
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_prediction_gpu.py tests/test_samples_gpu.py -x -q > gpurun_out/pred_tests.log 2>&1
tail -15 gpurun_out/pred_tests.log
