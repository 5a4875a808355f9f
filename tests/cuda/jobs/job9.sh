mkdir -p gpurun_out
{
echo "=== bf16 model test"
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q --tb=short 2>&1 | tail -30
echo "=== dbscan + plan cache tests"
timeout 900 python -m pytest tests/test_gpu_dbscan.py tests/test_gpu_plan_cache.py -x -q --tb=short 2>&1 | tail -15
echo "=== whole gpu suite"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== bench pipeline (grid dbscan)"
timeout 900 python bench.py --config pipeline --steps 3 --warmup 3 --trials 20
echo "=== sanitizer memcheck (API errors off)"
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --log-file gpurun_out/r2_sanitizer_memcheck_f32.log python tests/cuda/tc_diag.py 18 > gpurun_out/r2_san1.out 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck_f32.log; tail -1 gpurun_out/r2_san1.out | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --log-file gpurun_out/r2_sanitizer_memcheck_bf16.log python -m pytest tests/test_gpu_bf16.py -q -k "in_edge" > gpurun_out/r2_san2.out 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck_bf16.log; tail -2 gpurun_out/r2_san2.out
echo "=== sanitizer racecheck with the debug barrier beside out_ready"
GTB_EW_DEBUG_BAR=1 timeout 900 compute-sanitizer --tool racecheck --report-api-errors no --log-file gpurun_out/r2_sanitizer_racecheck_f32_debugbar.log python tests/cuda/tc_diag.py 17 > gpurun_out/r2_san3.out 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck_f32_debugbar.log
} > gpurun_out/r2_job9.log 2>&1
