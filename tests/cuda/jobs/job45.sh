mkdir -p gpurun_out
{
python -c "import os; print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))"
nvidia-smi topo -m | head -12
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 --multi independent 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('independent', d['ms_per_step'], d['e2e']['ms_per_step'])"
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/r2_n4_nccl.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('partitioned', d['ms_per_step'], d['e2e']['ms_per_step'])"
grep -E "NVLS|P2P|Channel|via|Connected|Using network|nranks" gpurun_out/r2_n4_nccl.err | head -20
} > gpurun_out/r2_job45.log 2>&1
