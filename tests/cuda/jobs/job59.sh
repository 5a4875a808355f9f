#!/bin/bash
# cell-list radius graph: parity (oracle, all-pairs walk), the DBSCAN tests over the refactored builder, timing
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_radius_losses.py tests/test_gpu_dbscan.py -x -q -m gpu > gpurun_out/job59_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/job59_tests.log
tail -5 gpurun_out/job59_tests.log
timeout 60 python tests/cuda/radius_time.py > gpurun_out/job59_time.log 2>&1
tail -2 gpurun_out/job59_time.log
