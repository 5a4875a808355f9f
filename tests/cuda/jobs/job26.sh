mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^frame" | tail -40
} > gpurun_out/r2_job26.log 2>&1
