"""Time of one EC training step (forward + BCE loss + backward) on the bench graph (run on the GPU box)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from gnn_tracking_b200.metrics.losses.ec import EdgeWeightBCELoss  # noqa: E402
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN  # noqa: E402

dims = sys.argv[1] if len(sys.argv) > 1 else "wide"
g = bench.relabel_by_phi(bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0))
dev = torch.device("cuda")
torch.manual_seed(0)
model = ECForGraphTCN(**bench.model_kwargs(dims)).to(dev)
x, ei, ea = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev)
y = (torch.rand(ei.size(1), device=dev) < 0.3)
loss_fn = EdgeWeightBCELoss()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)


def fwd():
    with torch.no_grad():
        return model.forward_tensors(x, ei, ea)["W"]


def train():
    opt.zero_grad(set_to_none=True)
    loss = loss_fn(w=model.forward_tensors(x, ei, ea)["W"], y=y)
    loss.backward()
    opt.step()
    return loss


for name, fn in (("forward", fwd), ("train step", train)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        fn()
    t.record()
    torch.cuda.synchronize()
    print(f"{dims} {name}: {s.elapsed_time(t) / 10:.2f} ms, peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
