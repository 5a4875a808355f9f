mkdir -p gpurun_out
{
timeout 600 python bench.py --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['serial_ms_per_step'])"
timeout 600 python bench.py --no-cpu --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['serial_ms_per_step'])"
} > gpurun_out/r2_job51.log 2>&1
