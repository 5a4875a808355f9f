mkdir -p gpurun_out
{
for c in 17 18 19; do timeout 120 python tests/cuda/tc_diag.py $c; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_perm'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20 | cut -c1-900
} > gpurun_out/r2_job32.log 2>&1
