"""``torch_cluster.radius_graph`` on the B200 library (the reference imports the un-vendored
torch_cluster op at metrics/losses/oc.py:7, metrics/losses/metric_learning.py:6).  The reference's
two losses do not need the edge list here (their sums are fused into the neighbour walk,
``gtb_radius_pair_sum_f32``); this is the same walk materialised for every other caller."""
from __future__ import annotations

import os

import torch
from torch import Tensor

from . import ops
from ._lib import check, lib


def radius_graph(x: Tensor, r: float, batch: Tensor | None = None, loop: bool = False, max_num_neighbors: int = 32,
                 flow: str = "source_to_target", num_workers: int = 1, method: str | None = None) -> Tensor:
    """Edges between all points within distance ``r`` (strict), signature of
    ``torch_cluster.radius_graph``.  ``flow="source_to_target"``: row 0 = neighbour, row 1 = centre;
    edges are grouped by centre with ascending neighbours; a centre with more than
    ``max_num_neighbors`` neighbours keeps the lowest indices (torch_cluster's subset is
    implementation-defined).  ``method``: "grid" (default: the uniform cell list shared with the DBSCAN
    trial, ``gtb_radius_graph_grid_*``) or "brute" (the all-pairs walk, ``gtb_radius_graph_*``; also chosen
    by ``GTB_RADIUS_BRUTE=1``); the edge lists are identical."""
    if method is None:
        method = "brute" if os.environ.get("GTB_RADIUS_BRUTE") == "1" else "grid"
    if method not in ("grid", "brute"):
        raise ValueError(method)
    if flow not in ("source_to_target", "target_to_source"):
        raise ValueError(flow)
    dev = ops.require_cuda(x)
    x = x.detach().to(torch.float32).contiguous()
    if x.dim() == 1:
        x = x.unsqueeze(1)
    n, d = x.shape
    if n == 0:
        return torch.empty((2, 0), dtype=torch.int64, device=dev)
    b = None if batch is None else batch.to(torch.int64).contiguous()
    st = ops.stream_ptr(dev)
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    bp = None if b is None else b.data_ptr()
    if method == "grid":
        ws_bytes = lib().gtb_radius_graph_grid_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with ops.on_device(dev):
            check(lib().gtb_radius_graph_grid_count_f32(x.data_ptr(), d, n, bp, float(r), int(max_num_neighbors), int(loop),
                                                        counts.data_ptr(), ws.data_ptr(), ws_bytes, st))
            incl = torch.cumsum(counts, 0, dtype=torch.int64)
            n_edges = int(incl[-1].item())  # sizes the output, as torch_cluster does
            offsets = (incl - counts).contiguous()
            edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
            check(lib().gtb_radius_graph_grid_fill_f32(x.data_ptr(), d, n, bp, float(r), int(max_num_neighbors), int(loop),
                                                       offsets.data_ptr(), edge_index.data_ptr(), n_edges, ws.data_ptr(),
                                                       ws_bytes, st))
        ops._count(7)  # bounding box, grid, keys, layout, count, fill (+ the radix sort)
        return edge_index if flow == "source_to_target" else edge_index.flip(0)
    check(lib().gtb_radius_graph_count_f32(x.data_ptr(), d, n, bp, float(r),
                                           int(max_num_neighbors), int(loop), counts.data_ptr(), st))
    incl = torch.cumsum(counts, 0, dtype=torch.int64)
    n_edges = int(incl[-1].item())  # sizes the output, as torch_cluster does
    offsets = (incl - counts).contiguous()
    edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
    check(lib().gtb_radius_graph_fill_f32(x.data_ptr(), d, n, None if b is None else b.data_ptr(), float(r),
                                          int(max_num_neighbors), int(loop), offsets.data_ptr(), edge_index.data_ptr(),
                                          n_edges, st))
    ops._count(2)
    return edge_index if flow == "source_to_target" else edge_index.flip(0)
