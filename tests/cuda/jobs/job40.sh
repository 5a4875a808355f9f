mkdir -p gpurun_out
timeout 300 python tests/cuda/save_hidden_dbg.py > gpurun_out/r2_job40.log 2>&1
