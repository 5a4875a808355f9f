mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_in_parity.py tests/test_gpu_fullsize.py tests/test_gpu_node_fused.py tests/test_gpu_atb.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_perm'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_wide.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1; echo rc=$?
} > gpurun_out/r2_job34.log 2>&1
