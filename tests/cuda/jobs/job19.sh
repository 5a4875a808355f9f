mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1500 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --graph uniform > gpurun_out/r2_bench_n2_uniform.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_n2_uniform.json
} > gpurun_out/r2_job19.log 2>&1
