mkdir -p gpurun_out
{
echo "=== pytest gpu (all)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== bench tcn_bf16"
timeout 900 python bench.py --config tcn_bf16 --steps 5 --warmup 3
echo "=== ncu launch list (EC forward)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_wide.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
echo "rc=$?"
echo "=== ncu full: fp32 edge kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_edge_ws -c 1 -f -o gpurun_out/r2_edge_ws_f32 python tests/cuda/tc_diag.py 20 > gpurun_out/r2_ncu_f32.log 2>&1
echo "rc=$?"
echo "=== ncu full: bf16 edge kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_edge_ws -c 1 -f -o gpurun_out/r2_edge_ws_bf16 python tests/cuda/bf16_edge_time.py > gpurun_out/r2_ncu_bf16.log 2>&1
echo "rc=$?"
echo "=== compute-sanitizer memcheck"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_f32.log python tests/cuda/tc_diag.py 18 > gpurun_out/r2_san1.out 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_f32.log
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_bf16.log python -m pytest tests/test_gpu_bf16.py -q -k "in_edge" > gpurun_out/r2_san2.out 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_bf16.log
echo "=== compute-sanitizer racecheck"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_f32.log python tests/cuda/tc_diag.py 17 > gpurun_out/r2_san3.out 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck_f32.log
} > gpurun_out/r2_job8.log 2>&1
