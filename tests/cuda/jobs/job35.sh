mkdir -p gpurun_out
{
timeout 600 ncu --set full --clock-control none --import-source on -k regex:in_node_ws -s 5 -c 1 -f -o gpurun_out/r2_node_ws python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_node_ws_ncu.log 2>&1; echo "rc=$?"
} > gpurun_out/r2_job35.log 2>&1
