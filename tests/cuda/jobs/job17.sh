mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_node_fused.py -x -q 2>&1 | grep -v "^frame\|^Search\|^CUDA kernel\|^For debugging\|^Compile with" | tail -25
timeout 600 python bench.py --steps 10 --warmup 3 | cut -c1-3000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_nodews.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_nodews.log 2>&1; echo rc=$?
} > gpurun_out/r2_job17.log 2>&1
