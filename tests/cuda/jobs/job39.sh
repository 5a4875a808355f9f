mkdir -p gpurun_out
{
for v in "X=1" "GTB_NO_SAVE_HIDDEN=1" "GTB_NO_ATB_TC=1" "GTB_NO_EDGE_WS=1"; do
echo "== $v"
env $v timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "test_in_layer_backward and auto" 2>&1 | grep -E "AssertionError:|passed|failed" | head -5
done
} > gpurun_out/r2_job39.log 2>&1
