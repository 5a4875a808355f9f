mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_node_fused.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_in_parity.py tests/test_gpu_fullsize.py tests/test_gpu_cabi.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu
GTB_NO_NODE_WS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu | cut -c1-200
} > gpurun_out/r2_job13.log 2>&1
