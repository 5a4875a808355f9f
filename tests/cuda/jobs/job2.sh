mkdir -p gpurun_out
{
echo "=== pytest with edge_ws"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== diag edge_ws prof"
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 14
GTB_NO_EDGE_WS=1 TC_PROF=1 timeout 300 python tests/cuda/tc_diag.py 14
} > gpurun_out/r2_job2.log 2>&1
