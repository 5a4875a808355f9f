mkdir -p gpurun_out
{
echo "=== diag"
for c in 8 17 18 19; do timeout 120 python tests/cuda/tc_diag.py $c; done
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20
echo "=== bf16 tests"
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -3
timeout 300 python tests/cuda/bf16_edge_time.py
echo "=== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
} > gpurun_out/r2_job11.log 2>&1
