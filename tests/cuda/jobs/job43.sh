mkdir -p gpurun_out
{
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; tail -c 300 gpurun_out/$2.json; }
run 29521 r2_bench_n8
run 29522 r2_bench_n8_uniform --graph uniform
run 29523 r2_bench_n8_config4 --total-scale 5
} > gpurun_out/r2_job43.log 2>&1
