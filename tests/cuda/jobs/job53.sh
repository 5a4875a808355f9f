mkdir -p gpurun_out
{
date
GTB_BENCH_GRAPH_MULTI=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_graph.json 2> gpurun_out/r2_bench_n2_graph.err; echo rc=$?
date
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n2_graph.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['parity']['max_err_over_scale'], d['config']['launch'])"
tail -3 gpurun_out/r2_bench_n2_graph.err | cut -c1-200
} > gpurun_out/r2_job53.log 2>&1
