mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_node_fused.py -x -q -k "test_ec_fused_stack" 2>&1 | grep -v "^frame\|^Search\|^CUDA kernel\|^For debugging\|^Compile with" | tail -15
timeout 600 python -m pytest tests/test_gpu_in_parity.py tests/test_gpu_fullsize.py tests/test_gpu_cabi.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 | cut -c1-2600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_wide.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1; echo rc=$?
} > gpurun_out/r2_job29.log 2>&1
