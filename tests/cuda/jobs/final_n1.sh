# Evidence run of the round on one B200: every file lands in gpurun_out/ and is copied to profiles/ by hand.
mkdir -p gpurun_out
O=gpurun_out
{
echo "=== tma probe"
for m in 0 1 2 3; do timeout 60 tests/cuda/tma_probe $m; done > $O/r2_tma_probe.log 2>&1; tail -4 $O/r2_tma_probe.log
echo "=== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r2_pytest_gpu_tail.log
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench lines"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_wide.json 2> $O/r2_bench_wide.err; tail -c 600 $O/r2_bench_wide.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference_cpu.json 2>/dev/null
timeout 900 python bench.py --dims default --steps 20 --warmup 5 > $O/r2_bench_default.json 2>/dev/null
timeout 900 python bench.py --graph uniform --steps 20 --warmup 5 --no-cpu > $O/r2_bench_uniform.json 2>/dev/null
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > $O/r2_bench_train.json 2>/dev/null
timeout 900 python bench.py --config tcn_bf16 --steps 5 --warmup 3 > $O/r2_bench_tcn_bf16.json 2>/dev/null
timeout 900 python bench.py --config pipeline --steps 3 --warmup 3 --trials 20 > $O/r2_bench_pipeline.json 2>/dev/null
GTB_NO_EDGE_WS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > $O/r2_bench_wide_generic_tiles.json 2>/dev/null
for f in default uniform train tcn_bf16 pipeline wide_generic_tiles reference_cpu; do echo "-- $f"; cut -c1-260 $O/r2_bench_$f.json; done
echo "=== stage profile"
EW_PROF=1 timeout 300 python tests/cuda/tc_diag.py 20 > $O/r2_edge_ws_stage_profile.txt 2>&1; cut -c1-300 $O/r2_edge_ws_stage_profile.txt
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_wide.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_bench.log 2>&1; echo rc=$?
echo "=== ncu full edge kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_edge_ws -c 1 -f -o $O/r2_edge_ws_f32 python tests/cuda/tc_diag.py 20 > $O/r2_ncu_f32.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_edge_ws -c 1 -f -o $O/r2_edge_ws_bf16 python tests/cuda/bf16_edge_time.py > $O/r2_ncu_bf16.log 2>&1; echo rc=$?
} > gpurun_out/r2_final_n1.log 2>&1
