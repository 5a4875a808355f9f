#!/usr/bin/env python
"""Times the loader-style end-to-end loop of bench.py (DevicePrefetcher, copy of step k + 1 under step k)
several times in a row, with per-step host timestamps: tells a slow GPU step from a stalled host."""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from gnn_tracking_b200.graph_store import DevicePrefetcher, GraphData  # noqa: E402
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN  # noqa: E402
from gnn_tracking_b200.plan import clear_plan_cache  # noqa: E402

dev = torch.device("cuda", 0)
g = bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0)
torch.manual_seed(0)
model = ECForGraphTCN(**bench.model_kwargs("wide")).to(dev)
hx, hei, hea = g["x"].pin_memory(), g["edge_index"].pin_memory(), g["edge_attr"].pin_memory()
hw = torch.empty(g["n_edges"], dtype=torch.float32).pin_memory()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
host_graph = GraphData(x=hx, edge_index=hei, edge_attr=hea)
do_flush = os.environ.get("NO_FLUSH") is None


def run(k, stamps=None):
    for data in DevicePrefetcher((host_graph for _ in range(k)), dev):
        if do_flush:
            flush.zero_()
        clear_plan_cache()
        with torch.no_grad():
            out = model.forward_tensors(data.x, data.edge_index, data.edge_attr)
        hw.copy_(out["W"], non_blocking=True)
        if stamps is not None:
            stamps.append(time.perf_counter())


run(5)
torch.cuda.synchronize()
for rep in range(4):
    stamps = []
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    run(20, stamps)
    t.record()
    torch.cuda.synchronize()
    host = [(b - a) * 1e3 for a, b in zip([t0] + stamps[:-1], stamps)]
    print(f"rep {rep}: {s.elapsed_time(t) / 20:.3f} ms/step (events)  host per step: " + " ".join(f"{h:.2f}" for h in host), flush=True)
# the same steps with resident inputs, host enqueue time only
x, ei, ea = hx.to(dev), hei.to(dev), hea.to(dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    clear_plan_cache()
    with torch.no_grad():
        model.forward_tensors(x, ei, ea)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"resident: host enqueue {1e3 * (t1 - t0) / 20:.3f} ms/step, drained after {1e3 * (t2 - t0) / 20:.3f} ms/step")
