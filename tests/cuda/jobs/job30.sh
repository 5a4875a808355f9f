mkdir -p gpurun_out
{
timeout 300 python tests/cuda/e2e_probe.py
NO_FLUSH=1 timeout 300 python tests/cuda/e2e_probe.py
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
} > gpurun_out/r2_job30.log 2>&1
