mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_plan_cache.py -x -q 2>&1 | grep -v "^frame" | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e6, d['gpu_launches'], d['roofline']['frac'], d['e2e'], d['parity'], d['config']['launch'])"
GTB_BENCH_NO_GRAPH=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['launch'])"
} > gpurun_out/r2_job46.log 2>&1
