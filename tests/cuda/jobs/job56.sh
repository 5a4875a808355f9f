mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graphs.py -x -q 2>&1 | tail -3 > gpurun_out/r2_job56.log 2>&1
