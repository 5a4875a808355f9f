mkdir -p gpurun_out
{
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_n$N.json
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --graph uniform > gpurun_out/r2_bench_n2_uniform.json 2>/dev/null; echo "rc=$?"
fi
} > gpurun_out/r2_job44_$1.log 2>&1
