mkdir -p gpurun_out
{
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo rc=$?
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['parity']['max_err_over_scale'], d['config']['launch'])"
} > gpurun_out/r2_job52.log 2>&1
