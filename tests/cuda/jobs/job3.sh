mkdir -p gpurun_out
{
echo "=== diag edge_ws v2"
for c in 8 17 18 19; do timeout 120 python tests/cuda/tc_diag.py $c; done
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 14
echo "=== pytest with edge_ws"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench edge_ws"
timeout 600 python bench.py --steps 10 --warmup 3
} > gpurun_out/r2_job3.log 2>&1
