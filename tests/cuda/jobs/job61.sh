#!/bin/bash
# final check of the tree: full GPU suite, smoke, one default bench line
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_tail.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
timeout 300 python bench.py > gpurun_out/r2_bench_final.json
tail -c 1500 gpurun_out/r2_bench_final.json
} > gpurun_out/r2_job61.log 2>&1
tail -12 gpurun_out/r2_job61.log
