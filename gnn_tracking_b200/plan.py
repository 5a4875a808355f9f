"""Graph plan: the destination-sorted view of ``edge_index`` that every layer of a
stack shares.  Replaces the per-call ``index_select`` / ``scatter_add_`` derivation
inside PyG's ``MessagePassing.propagate`` (reference
models/interaction_network.py:67)."""
from __future__ import annotations

import weakref
from dataclasses import dataclass

import torch
from torch import Tensor

from . import ops
from ._lib import check, lib


@dataclass
class GraphPlan:
    n_nodes: int
    n_edges: int
    perm: Tensor        # int32 [E]  stable argsort of edge_index[1]
    rowptr: Tensor      # int32 [N+1]
    src_sorted: Tensor  # int32 [E]  edge_index[0][perm]
    dst_sorted: Tensor  # int32 [E]  edge_index[1][perm]
    status: Tensor      # int32 [1]  nonzero: an index was out of range

    def validate(self) -> None:
        """Host-synchronising range check (the reference would raise an
        IndexError inside index_select)."""
        if int(self.status.item()) != 0:
            raise IndexError("edge_index contains node indices outside [0, num_nodes)")

    def filtered(self, keep: Tensor) -> tuple["GraphPlan", Tensor, Tensor]:
        """Plan of ``Data.edge_subgraph(keep)`` (reference
        models/track_condensation_networks.py:251-252) by stream compaction of this
        plan -- the sorted order survives filtering, no re-sort.  Returns the plan and
        ``new_id`` (int32 [E]: position of each kept edge in the compacted edge list,
        -1 for dropped edges) and ``kept_ids`` (int32 [E']: original id of each kept edge).  One host sync to learn the kept-edge count (the
        reference's boolean indexing syncs at the same place)."""
        dev = ops.require_cuda(keep)
        if keep.dtype != torch.bool or keep.numel() != self.n_edges:
            raise ValueError("keep must be a bool mask over the edges")
        keep8 = keep.contiguous().view(torch.uint8)
        e, n = self.n_edges, self.n_nodes
        i32 = dict(dtype=torch.int32, device=dev)
        new_id = torch.empty(e, **i32)
        kept = torch.empty(e, **i32)
        perm = torch.empty(e, **i32)
        src = torch.empty(e, **i32)
        dst = torch.empty(e, **i32)
        rowptr = torch.empty(n + 1, **i32)
        n_kept = torch.zeros(1, **i32)
        ws_bytes = lib().gtb_plan_filter_workspace_bytes(n, e)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().gtb_plan_filter(keep8.data_ptr(), n, e, self.perm.data_ptr(), self.src_sorted.data_ptr(),
                                    self.dst_sorted.data_ptr(), new_id.data_ptr(), kept.data_ptr(), perm.data_ptr(),
                                    rowptr.data_ptr(), src.data_ptr(), dst.data_ptr(), n_kept.data_ptr(),
                                    ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
        ops._count(6)
        k = int(n_kept.item())
        sub = GraphPlan(n, k, perm[:k], rowptr, src[:k], dst[:k], torch.zeros(1, **i32))
        return sub, new_id, kept[:k]


def build_plan(edge_index: Tensor, n_nodes: int) -> GraphPlan:
    dev = ops.require_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise TypeError("edge_index must be an int64 tensor of shape [2, E]")
    ei = edge_index.contiguous()
    e = ei.size(1)
    i32 = dict(dtype=torch.int32, device=dev)
    perm = torch.empty(e, **i32)
    src = torch.empty(e, **i32)
    dst = torch.empty(e, **i32)
    rowptr = torch.empty(n_nodes + 1, **i32)
    status = torch.empty(1, **i32)
    ws_bytes = lib().gtb_plan_workspace_bytes(n_nodes, e)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib().gtb_plan_build(ei.data_ptr(), n_nodes, e, perm.data_ptr(), rowptr.data_ptr(), src.data_ptr(),
                               dst.data_ptr(), status.data_ptr(), ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
    ops._count(6)  # keys, radix-sort passes, rowptr
    return GraphPlan(n_nodes, e, perm, rowptr, src, dst, status)


# One plan per live edge_index tensor: all layers of a stack (and repeated forwards over the
# same graph object) share it.  Keyed on identity + version so in-place edits invalidate it.
_CACHE: dict[int, tuple[weakref.ref, int, int, GraphPlan]] = {}


def get_plan(edge_index: Tensor, n_nodes: int) -> GraphPlan:
    key = id(edge_index)
    hit = _CACHE.get(key)
    if hit is not None:
        ref, version, n, plan = hit
        if ref() is edge_index and version == edge_index._version and n == n_nodes \
                and plan.n_edges == edge_index.size(1):
            return plan
    plan = build_plan(edge_index, n_nodes)
    if len(_CACHE) > 64:
        for k in [k for k, v in _CACHE.items() if v[0]() is None]:
            del _CACHE[k]
        if len(_CACHE) > 64:
            _CACHE.clear()
    _CACHE[key] = (weakref.ref(edge_index), edge_index._version, n_nodes, plan)
    return plan


def adopt_plan(edge_index: Tensor, n_nodes: int, plan: GraphPlan) -> None:
    """Register a plan computed elsewhere (``graph_store.read_graph``: stored next to the graph by the
    offline writer) for this ``edge_index`` tensor: ``get_plan`` returns it without sorting."""
    if plan.n_edges != edge_index.size(1) or plan.n_nodes != n_nodes:
        raise ValueError("plan does not match the graph")
    _CACHE[id(edge_index)] = (weakref.ref(edge_index), edge_index._version, n_nodes, plan)


def clear_plan_cache() -> None:
    _CACHE.clear()
