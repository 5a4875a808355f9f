mkdir -p gpurun_out
{
timeout 600 python tests/cuda/host_overhead.py
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
} > gpurun_out/r2_job18.log 2>&1
