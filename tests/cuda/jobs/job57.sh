mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_in_parity.py -x -q -k "batch_norm or skip2" 2>&1 | tail -8 > gpurun_out/r2_job57.log 2>&1
