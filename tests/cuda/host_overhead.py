"""Host-side cost of one EC forward (run on the GPU box): wall time to ENQUEUE a step vs the GPU time
of the step, and the Python hot spots (cProfile)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN  # noqa: E402
from gnn_tracking_b200.plan import clear_plan_cache  # noqa: E402

g = bench.relabel_by_phi(bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0))
dev = torch.device("cuda")
torch.manual_seed(0)
model = ECForGraphTCN(**bench.model_kwargs("wide")).to(dev)
x, ei, ea = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev)


def step():
    clear_plan_cache()
    with torch.no_grad():
        return model.forward_tensors(x, ei, ea)


for _ in range(5):
    step()
torch.cuda.synchronize()
K = 30
t0 = time.perf_counter()
for _ in range(K):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, enqueue + drain {1e3 * (t2 - t0) / K:.3f} ms/step")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
