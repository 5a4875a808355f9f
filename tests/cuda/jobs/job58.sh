mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_tail.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
} > gpurun_out/r2_job58.log 2>&1
