mkdir -p gpurun_out
{
date
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo rc=$?
date
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['parity']['max_err_over_scale'], d['config']['launch'])"
tail -3 gpurun_out/r2_bench_n8.err | cut -c1-200
} > gpurun_out/r2_job54.log 2>&1
