mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_graph_store.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['layer_ms'], d['roofline']['layer_frac'], d['e2e'])"
} > gpurun_out/r2_job37.log 2>&1
