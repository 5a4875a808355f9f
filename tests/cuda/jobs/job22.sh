mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_atb.py -x -q 2>&1 | grep -v "^frame\|^Search\|^CUDA kernel\|^For debugging\|^Compile with" | tail -25
timeout 300 python tests/cuda/atb_time.py
GTB_NO_ATB_TC=1 timeout 300 python tests/cuda/atb_time.py
} > gpurun_out/r2_job22.log 2>&1
