mkdir -p gpurun_out
{
echo "=== bf16 tests"
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -25
echo "=== bench tcn_bf16"
timeout 900 python bench.py --config tcn_bf16 --steps 5 --warmup 3
echo "=== bench pipeline"
timeout 900 python bench.py --config pipeline --steps 3 --warmup 3 --trials 10
echo "=== bench train"
timeout 900 python bench.py --mode train --steps 5 --warmup 3
echo "=== bench uniform control"
timeout 900 python bench.py --graph uniform --steps 10 --warmup 3 --no-cpu
} > gpurun_out/r2_job7.log 2>&1
