"""One warm EC training step under ncu's launch list (run: ncu --metrics gpu__time_duration.sum ... python this)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from gnn_tracking_b200.metrics.losses.ec import EdgeWeightBCELoss  # noqa: E402
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN  # noqa: E402

g = bench.relabel_by_phi(bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0))
dev = torch.device("cuda")
torch.manual_seed(0)
model = ECForGraphTCN(**bench.model_kwargs("wide")).to(dev)
x, ei, ea = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev)
y = (torch.rand(ei.size(1), device=dev) < 0.3)
loss_fn = EdgeWeightBCELoss()
for _ in range(2):
    model.zero_grad(set_to_none=True)
    loss_fn(w=model.forward_tensors(x, ei, ea)["W"], y=y).backward()
torch.cuda.synchronize()
