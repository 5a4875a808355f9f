mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_losses_tcn.py tests/test_gpu_radius_losses.py tests/test_gpu_cabi.py -x -q 2>&1 | tail -4
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu | cut -c1-330
GTB_NO_SAVE_HIDDEN=1 timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu | cut -c1-230
} > gpurun_out/r2_job38.log 2>&1
