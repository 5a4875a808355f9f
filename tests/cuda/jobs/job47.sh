mkdir -p gpurun_out
{
GTB_BENCH_GRAPH_MULTI=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2_n2_graph.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph', d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d.get('parity'), d['config']['launch'])"
tail -5 gpurun_out/r2_n2_graph.err | cut -c1-300
} > gpurun_out/r2_job47.log 2>&1
