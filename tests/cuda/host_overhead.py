"""Host enqueue time of one EC forward (no synchronisation inside the loop) next to its GPU time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
from gnn_tracking_b200.plan import clear_plan_cache

g = bench.make_graph(bench.N_NODES, bench.N_EDGES)
torch.manual_seed(0)
m = ECForGraphTCN(**bench.model_kwargs("wide")).cuda()
x, ei, ea = g["x"].cuda(), g["edge_index"].cuda(), g["edge_attr"].cuda()
for label, env in (("fused", None), ("GTB_NO_NODE_WS+GTB_NO_ENC_WS", "1")):
    if env:
        os.environ["GTB_NO_NODE_WS"] = "1"; os.environ["GTB_NO_ENC_WS"] = "1"
    with torch.no_grad():
        for _ in range(5):
            clear_plan_cache(); m.forward_tensors(x, ei, ea)
        torch.cuda.synchronize()
        for reps in (1, 20):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); s.record()
            for _ in range(reps):
                clear_plan_cache(); m.forward_tensors(x, ei, ea)
            e.record(); t1 = time.perf_counter()
            torch.cuda.synchronize()
            print(f"{label}: reps={reps} host enqueue {1e3 * (t1 - t0) / reps:.3f} ms/step, gpu {s.elapsed_time(e) / reps:.3f} ms/step")
