mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_atb.py tests/test_gpu_backward.py -x -q 2>&1 | tail -3
timeout 300 python tests/cuda/atb_time.py | tail -3
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu | cut -c1-200
} > gpurun_out/r2_job48.log 2>&1
