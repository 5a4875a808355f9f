mkdir -p gpurun_out
{
S="compute-sanitizer --report-api-errors no"
timeout 600 $S --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_node_ws.log python -m pytest tests/test_gpu_node_fused.py -q -x -k "test_node_fused_vs_float64 and (300 or 127 or 4096)" 2>&1 | tail -2
timeout 600 $S --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_enc_ws.log python -m pytest tests/test_gpu_node_fused.py -q -x -k "test_edge_encoder_vs_float64 and (300 or 127)" 2>&1 | tail -2
timeout 600 $S --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_atb_tc.log python -m pytest tests/test_gpu_atb.py -q -x -k "test_rows_atb_tensor_core and 5000" 2>&1 | tail -2
timeout 600 $S --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_node_ws.log python -m pytest tests/test_gpu_node_fused.py -q -x -k "test_node_fused_vs_float64 and 300 and obj+proj" 2>&1 | tail -2
timeout 600 $S --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_enc_ws.log python -m pytest tests/test_gpu_node_fused.py -q -x -k "test_edge_encoder_vs_float64 and 300" 2>&1 | tail -2
timeout 600 $S --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_atb_tc.log python -m pytest tests/test_gpu_atb.py -q -x -k "test_rows_atb_tensor_core and 5000" 2>&1 | tail -2
for f in gpurun_out/r2_sanitizer_*node_ws.log gpurun_out/r2_sanitizer_*enc_ws.log gpurun_out/r2_sanitizer_*atb_tc.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $f; done
} > gpurun_out/r2_job24.log 2>&1
