#!/usr/bin/env python
"""Times the bf16 IN edge kernel (128 / 128 / 128) on a 100k-node / 1M-edge TrackML-shaped graph, L2 flushed."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from gnn_tracking_b200 import ops  # noqa: E402
from gnn_tracking_b200.plan import build_plan  # noqa: E402

dev = torch.device("cuda", 0)
g = bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0)
n, e = g["n_nodes"], g["n_edges"]
plan = build_plan(g["edge_index"].to(dev), n)
gen = torch.Generator().manual_seed(0)
bf = torch.bfloat16
e_in = torch.randn(e, 128, generator=gen).to(bf).to(dev)
p_i = torch.randn(n, 128, generator=gen).to(bf).to(dev)
p_j = torch.randn(n, 128, generator=gen).to(bf).to(dev)
ws = [(torch.randn(128, 128, generator=gen) / 128 ** 0.5).to(dev) for _ in range(3)]
bs = [(torch.randn(128, generator=gen) * 0.1).to(dev) for _ in range(3)]
packed = ops.pack_in_edge_bf16(ws, bs)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for mode, kw in (("sorted", {}), ("perm", dict(e_index=plan.perm, out_index=plan.perm))):
    ts = []
    for i in range(13):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.in_edge_bf16(e_in, p_i, p_j, plan.src_sorted, plan.dst_sorted, packed, n, relu_e=True, **kw)
        t.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(t))
    ms = statistics.mean(ts)  # includes the zero-fill of the fp32 aggregate [N, 128] (51 MB)
    alg = e * (16 + 256 + 256) + n * 2 * (128 + 128)
    print(f"bf16 edge kernel ({mode}): {ms * 1e3:.1f} us per launch (+ aggregate zero-fill), {alg / ms / 1e6:.0f} GB/s algorithmic")
