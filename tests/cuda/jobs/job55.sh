mkdir -p gpurun_out
{
date
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo rc=$?
date
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['parity']['max_err_over_scale'], d['config']['launch'][:40])"
} > gpurun_out/r2_job55.log 2>&1
