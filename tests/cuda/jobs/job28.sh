mkdir -p gpurun_out
{
timeout 600 python tests/cuda/determinism_probe.py
GTB_NO_NODE_WS=1 GTB_NO_ENC_WS=1 timeout 600 python tests/cuda/determinism_probe.py
} > gpurun_out/r2_job28.log 2>&1
