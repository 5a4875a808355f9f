mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_plan_cache.py -x -q 2>&1 | tail -40
} > gpurun_out/r2_job27.log 2>&1
