mkdir -p gpurun_out
{
timeout 300 python tests/cuda/head_only.py
S="compute-sanitizer --report-api-errors no"
timeout 600 $S --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_head_ws.log python tests/cuda/head_only.py | tail -1
timeout 600 $S --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_head_ws.log python tests/cuda/head_only.py | tail -1
grep -E "SUMMARY" gpurun_out/r2_sanitizer_racecheck_head_ws.log gpurun_out/r2_sanitizer_memcheck_head_ws.log
grep -E "Error: Race" gpurun_out/r2_sanitizer_racecheck_head_ws.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head
} > gpurun_out/r2_job42.log 2>&1
