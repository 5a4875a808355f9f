mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_losses_tcn.py tests/test_gpu_radius_losses.py tests/test_gpu_cabi.py -q 2>&1 | grep -E "AssertionError|passed|failed|FAILED" | head
} > gpurun_out/r2_job41.log 2>&1
