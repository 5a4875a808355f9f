mkdir -p gpurun_out
{
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_nodews.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_nodews.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:in_node_ws -s 3 -c 1 -f -o gpurun_out/r2_node_ws python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_nodews_full.log 2>&1; echo rc=$?
} > gpurun_out/r2_job14.log 2>&1
