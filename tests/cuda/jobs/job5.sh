mkdir -p gpurun_out
{
echo "=== diag edge_ws v2c"
for c in 8 17 18 19; do timeout 120 python tests/cuda/tc_diag.py $c; done
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20
echo "=== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
} > gpurun_out/r2_job5.log 2>&1
