mkdir -p gpurun_out
{
for c in 17 18; do timeout 120 python tests/cuda/tc_diag.py $c; done
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -3
timeout 300 python tests/cuda/bf16_edge_time.py
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
} > gpurun_out/r2_job12.log 2>&1
