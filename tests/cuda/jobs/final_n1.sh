# Evidence run of the round on one B200: every file lands in gpurun_out/ and is copied to profiles/ by hand.
mkdir -p gpurun_out
O=gpurun_out
{
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r2_pytest_gpu_tail.log
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench lines"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_wide.json 2> $O/r2_bench_wide.err; tail -c 900 $O/r2_bench_wide.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference_cpu.json 2>/dev/null
timeout 900 python bench.py --dims default --steps 20 --warmup 5 > $O/r2_bench_default.json 2>/dev/null
timeout 900 python bench.py --graph uniform --steps 20 --warmup 5 --no-cpu > $O/r2_bench_uniform.json 2>/dev/null
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > $O/r2_bench_train.json 2>/dev/null
timeout 900 python bench.py --config tcn_bf16 --steps 5 --warmup 3 > $O/r2_bench_tcn_bf16.json 2>/dev/null
timeout 900 python bench.py --config pipeline --steps 3 --warmup 3 --trials 20 > $O/r2_bench_pipeline.json 2>/dev/null
GTB_NO_EDGE_WS=1 GTB_NO_NODE_WS=1 GTB_NO_ENC_WS=1 GTB_NO_HEAD_WS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > $O/r2_bench_wide_generic_tiles.json 2>/dev/null
for f in default uniform train tcn_bf16 pipeline wide_generic_tiles reference_cpu; do echo "-- $f"; cut -c1-260 $O/r2_bench_$f.json; done
echo "=== stage profile"
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20 > $O/r2_edge_ws_stage_profile.txt 2>&1; cut -c1-300 $O/r2_edge_ws_stage_profile.txt
echo "=== launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_wide.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_bench.log 2>&1; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > $O/r2_ncu_train.log 2>&1; echo rc=$?
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_edge_ws -c 1 -f -o $O/r2_edge_ws_f32 python tests/cuda/tc_diag.py 20 > $O/r2_ncu_f32.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_edge_ws -c 1 -f -o $O/r2_edge_ws_bf16 python tests/cuda/bf16_edge_time.py > $O/r2_ncu_bf16.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:in_node_ws -s 5 -c 1 -f -o $O/r2_node_ws python bench.py --steps 1 --warmup 3 --no-cpu > $O/r2_node_ws_ncu.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_encoder_ws -s 1 -c 1 -f -o $O/r2_enc_ws python bench.py --steps 1 --warmup 3 --no-cpu > $O/r2_enc_ws_ncu.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ec_head_ws -s 1 -c 1 -f -o $O/r2_head_ws python bench.py --steps 1 --warmup 3 --no-cpu > $O/r2_head_ws_ncu.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_atb_tc -c 1 -f -o $O/r2_atb_tc python tests/cuda/atb_time.py > $O/r2_atb_tc_ncu.log 2>&1; echo rc=$?
echo "=== sanitizer (head kernel, SAVE variant of the edge kernel)"
S="compute-sanitizer --report-api-errors no"
timeout 600 $S --tool memcheck --log-file $O/r2_sanitizer_memcheck_head_ws.log python -m pytest tests/test_gpu_node_fused.py -q -x -k "test_ec_fused_stack and skip1-3" 2>&1 | tail -1
timeout 600 $S --tool racecheck --log-file $O/r2_sanitizer_racecheck_head_ws.log python -m pytest tests/test_gpu_node_fused.py -q -x -k "test_ec_fused_stack and skip1-1" 2>&1 | tail -1
timeout 600 $S --tool memcheck --log-file $O/r2_sanitizer_memcheck_train_step.log python -m pytest tests/test_gpu_backward.py -q -x -k "test_edge_classifier_training_step_gradients and auto" 2>&1 | tail -1
for f in $O/r2_sanitizer_*head_ws.log $O/r2_sanitizer_memcheck_train_step.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $f; done
} > gpurun_out/r2_final_n1.log 2>&1
