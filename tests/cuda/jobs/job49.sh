mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_node_fused.py tests/test_gpu_fullsize.py tests/test_gpu_graphs.py -x -q 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_wide.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1; echo rc=$?
GTB_BENCH_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_wide_eager.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench2.log 2>&1; echo rc=$?
} > gpurun_out/r2_job49.log 2>&1
