mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -4
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu | cut -c1-700
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_train.log 2>&1; echo rc=$?
} > gpurun_out/r2_job23.log 2>&1
