mkdir -p gpurun_out
{
for k in in_node_ws:r2_node_ws:4 edge_encoder_ws:r2_enc_ws:1 ec_head_ws:r2_head_ws:1; do
  IFS=: read re out skip <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o gpurun_out/$out python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${out}_ncu.log 2>&1; echo "$out rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_atb_tc -c 1 -f -o gpurun_out/r2_atb_tc python tests/cuda/atb_time.py > gpurun_out/r2_atb_tc_ncu.log 2>&1; echo "atb rc=$?"
} > gpurun_out/r2_job31.log 2>&1
