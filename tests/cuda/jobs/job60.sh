#!/bin/bash
# cell-list pair sums: losses tests over both searches, grid == brute, timing
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_radius_losses.py tests/test_gpu_losses_tcn.py -x -q -m gpu > gpurun_out/job60_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/job60_tests.log
tail -8 gpurun_out/job60_tests.log
timeout 60 python tests/cuda/pair_sum_time.py > gpurun_out/job60_time.log 2>&1
tail -2 gpurun_out/job60_time.log
