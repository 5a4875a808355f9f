mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_radius_losses.py tests/test_gpu_losses_tcn.py -x -q 2>&1 | tail -4
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_train.json; cut -c1-200 gpurun_out/r2_bench_train.json
} > gpurun_out/r2_job25.log 2>&1
