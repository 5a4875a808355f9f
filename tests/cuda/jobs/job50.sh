mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_tail.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/r2_bench_wide.json 2> gpurun_out/r2_bench_wide.err; tail -c 1500 gpurun_out/r2_bench_wide.json; tail -3 gpurun_out/r2_bench_wide.err
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2_bench_train.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_train.json
} > gpurun_out/r2_job50.log 2>&1
