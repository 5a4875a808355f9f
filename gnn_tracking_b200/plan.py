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
        IndexError inside index_select).  Without it the check is deferred: see ``_poll_status``."""
        _forget_status(self.status)  # settled here: the deferred check must not raise it a second time
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


def prune_orphans(plan: GraphPlan) -> tuple[GraphPlan, Tensor, Tensor]:
    """Plan of the graph without its edge-less nodes, the others relabelled in increasing order (reference
    models/track_condensation_networks.py:254-259), derived from ``plan`` without sorting again: the
    relabelling is monotone, so ``perm`` stays valid, the endpoints are relabelled and ``rowptr`` compacted
    (``gtb_plan_prune_orphans``).  Returns ``(plan', node_ids int32 [N'], new_id int32 [N])``: original id of
    every surviving node, and new id of every node (-1 for an orphan).  One host sync to learn N' (the
    reference's ``unique`` syncs at the same place: the outputs' shapes depend on it)."""
    dev = plan.perm.device
    n, e = plan.n_nodes, plan.n_edges
    i32 = dict(dtype=torch.int32, device=dev)
    new_id = torch.empty(n, **i32)
    node_ids = torch.empty(n, **i32)
    rowptr = torch.empty(n + 1, **i32)
    src = torch.empty(e, **i32)
    dst = torch.empty(e, **i32)
    n_kept = torch.zeros(1, **i32)
    ws_bytes = lib().gtb_plan_prune_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with ops.on_device(dev):
        check(lib().gtb_plan_prune_orphans(n, e, plan.rowptr.data_ptr(), plan.src_sorted.data_ptr(), plan.dst_sorted.data_ptr(),
                                           new_id.data_ptr(), node_ids.data_ptr(), rowptr.data_ptr(), src.data_ptr(),
                                           dst.data_ptr(), n_kept.data_ptr(), ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
    ops._count(4)
    k = int(n_kept.item())
    return GraphPlan(k, e, plan.perm, rowptr[:k + 1], src, dst, torch.zeros(1, **i32)), node_ids[:k], new_id


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
    with ops.on_device(dev):
        check(lib().gtb_plan_build(ei.data_ptr(), n_nodes, e, perm.data_ptr(), rowptr.data_ptr(), src.data_ptr(),
                                   dst.data_ptr(), status.data_ptr(), ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
    ops._count(6)  # keys, radix-sort passes, rowptr
    _defer_status_check(status)
    return GraphPlan(n_nodes, e, perm, rowptr, src, dst, status)


# One plan per live edge_index tensor: all layers of a stack (and repeated forwards over the
# same graph object) share it.  Keyed on identity + version so in-place edits invalidate it; the
# entry dies with the tensor (weakref callback), so a long epoch through ``GraphLoader`` does not
# keep stale graphs resident through their plans.
_CACHE: dict[int, tuple[weakref.ref, int | None, int, GraphPlan]] = {}


def _version(edge_index: Tensor) -> int | None:
    """Inference tensors (``torch.inference_mode()``: Lightning's validation / test / predict
    loops) do not track a version counter and cannot be mutated in place outside inference mode:
    identity + shape is enough for them."""
    return None if edge_index.is_inference() else edge_index._version


def _remember(edge_index: Tensor, n_nodes: int, plan: GraphPlan) -> None:
    key = id(edge_index)

    def _drop(ref, key=key):
        hit = _CACHE.get(key)
        if hit is not None and hit[0] is ref:
            del _CACHE[key]

    _CACHE[key] = (weakref.ref(edge_index, _drop), _version(edge_index), n_nodes, plan)


def get_plan(edge_index: Tensor, n_nodes: int) -> GraphPlan:
    if not torch.cuda.is_current_stream_capturing():
        _poll_status()
    hit = _CACHE.get(id(edge_index))
    if hit is not None:
        ref, version, n, plan = hit
        if ref() is edge_index and version == _version(edge_index) and n == n_nodes \
                and plan.n_edges == edge_index.size(1):
            return plan
    plan = build_plan(edge_index, n_nodes)
    _remember(edge_index, n_nodes, plan)
    return plan


def adopt_plan(edge_index: Tensor, n_nodes: int, plan: GraphPlan) -> None:
    """Register a plan computed elsewhere (``graph_store.read_graph``: stored next to the graph by the
    offline writer) for this ``edge_index`` tensor: ``get_plan`` returns it without sorting."""
    if plan.n_edges != edge_index.size(1) or plan.n_nodes != n_nodes:
        raise ValueError("plan does not match the graph")
    _remember(edge_index, n_nodes, plan)


def clear_plan_cache() -> None:
    _CACHE.clear()


# ---- deferred range check.  ``gtb_plan_build`` clamps node ids into [0, N) (the kernels that walk the
# plan can never leave their tables) and raises ``status``; the flag travels to a pinned host slot
# behind the build and is looked at -- without synchronising -- by the next plan calls, which raise
# the IndexError the reference's index_select would have raised (at the latest ``GraphPlan.validate``).
_STATUS_SLOTS = 32
_status_host: Tensor | None = None
_status_pending: list[tuple[torch.cuda.Event, int, int]] = []  # (copy done, host slot, id of the status tensor)
_status_next = 0


def _defer_status_check(status: Tensor) -> None:
    global _status_host, _status_next
    if torch.cuda.is_current_stream_capturing():
        return  # inside a CUDA graph (graphs.CapturedForward): the flag stays on the device, ``validate`` reads it
    if _status_host is None:
        _status_host = torch.zeros(_STATUS_SLOTS, dtype=torch.int32).pin_memory()
    if len(_status_pending) >= _STATUS_SLOTS:  # ring full: settle the oldest check now
        ev, slot, _ = _status_pending.pop(0)
        ev.synchronize()
        _raise_if_set(slot)
    slot = _status_next
    _status_next = (_status_next + 1) % _STATUS_SLOTS
    _status_host[slot:slot + 1].copy_(status, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(status.device))
    _status_pending.append((ev, slot, id(status)))


def _raise_if_set(slot: int) -> None:
    if int(_status_host[slot]) != 0:
        _status_host[slot] = 0
        raise IndexError("edge_index contains node indices outside [0, num_nodes)")


def _forget_status(status: Tensor) -> None:
    for i, (ev, slot, key) in enumerate(_status_pending):
        if key == id(status):
            ev.synchronize()
            _status_host[slot] = 0
            del _status_pending[i]
            return


def _poll_status() -> None:
    while _status_pending and _status_pending[0][0].query():
        _, slot, _ = _status_pending.pop(0)
        _raise_if_set(slot)
