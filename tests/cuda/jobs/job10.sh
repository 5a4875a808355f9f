mkdir -p gpurun_out
{
echo "=== 2-GPU tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -25
echo "=== bf16 model test"
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q --tb=short 2>&1 | tail -8
echo "=== bench N=2 partitioned (trackml)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3
echo "=== bench N=2 partitioned (uniform control)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --graph uniform
echo "=== bench pipeline"
timeout 900 python bench.py --config pipeline --steps 3 --warmup 3 --trials 20
} > gpurun_out/r2_job10.log 2>&1
