mkdir -p gpurun_out
{
echo "=== diag edge_ws v2b"
for c in 8 17 18 19; do timeout 120 python tests/cuda/tc_diag.py $c; done
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20
timeout 300 python tests/cuda/tc_diag.py 14
echo "=== e2e probe (edge_ws)"
timeout 300 python tests/cuda/e2e_probe.py
echo "=== e2e probe (generic)"
GTB_NO_EDGE_WS=1 timeout 300 python tests/cuda/e2e_probe.py
echo "=== pytest"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
} > gpurun_out/r2_job4.log 2>&1
