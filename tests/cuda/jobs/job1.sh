mkdir -p gpurun_out
{
for m in 0 1 2 3; do timeout 60 tests/cuda/tma_probe $m; echo "rc=$?"; done
echo "=== pytest (generic tiles only, sorted-order python path)"
GTB_NO_EDGE_WS=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== diag edge_ws"
for c in 8 17 18 19; do timeout 120 python tests/cuda/tc_diag.py $c; done
timeout 300 python tests/cuda/tc_diag.py 14
timeout 300 python tests/cuda/tc_diag.py 20
echo "=== pytest with edge_ws"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== bench generic"
GTB_NO_EDGE_WS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
echo "=== bench edge_ws"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
} > gpurun_out/r2_job1.log 2>&1
