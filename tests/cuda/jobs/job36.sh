mkdir -p gpurun_out
{
timeout 900 python bench.py --config tcn_bf16 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_tcn_bf16.json 2> gpurun_out/r2_bench_tcn_bf16.err; cut -c1-600 gpurun_out/r2_bench_tcn_bf16.json; tail -3 gpurun_out/r2_bench_tcn_bf16.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_tcn_bf16.csv python bench.py --config tcn_bf16 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_tcn.log 2>&1; echo rc=$?
} > gpurun_out/r2_job36.log 2>&1
