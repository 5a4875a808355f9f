"""Node-partitioned graphs across the GPUs of one box.

The reference has no multi-GPU path at all (single process, ``batch_size=1`` graph per step,
reference utils/loading.py:235); its only related idea is cutting an event into overlapping
phi sectors that are trained as independent graphs (preprocessing/point_cloud_builder.py:242-327).
Here ONE graph is sharded instead:

* every rank owns a contiguous range of node ids (callers relabel nodes by phi first, so that
  the tight ``phi_slope`` cut of the graph builder keeps most edges inside one range);
* an edge belongs to the owner of its DESTINATION: the sum aggregation and the node MLP
  (reference models/interaction_network.py:92-103) stay local and need no reduction;
* the sources a rank's edges reference but does not own are its HALO.  Per Interaction-Network
  layer the owners send the halo rows of ONE per-node table (the pre-projected source block of
  the relational model's first Linear, 4 * H bytes per row) -- a single all-to-all-v, issued as
  one batch of point-to-point sends / receives (NCCL groups them into one launch; the same code
  runs on gloo in the CPU tests).  Nothing else of the path communicates.

``partition_graph`` is host logic on index tensors (built once per graph, like the plan);
``HaloExchange.extend`` is the only call inside the forward.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist
from torch import Tensor


@dataclass
class GraphShard:
    """Rank ``rank``'s part of a graph.  Local node numbering: owned nodes first
    (``global id - node_lo``), then the halo nodes in ascending global id."""
    rank: int
    world: int
    node_lo: int
    node_hi: int
    halo_ids: Tensor            # int64 [n_halo] global ids of the halo nodes (sorted)
    edge_ids: Tensor            # int64 [E_p] global ids of the owned edges, original order kept
    edge_index: Tensor          # int64 [2, E_p] local numbering: row 0 in [0, n_owned + n_halo), row 1 in [0, n_owned)
    send_idx: Tensor            # int32 [sum(send_counts)] LOCAL ids of owned rows to send, grouped by destination rank
    send_counts: list[int] = field(default_factory=list)
    recv_counts: list[int] = field(default_factory=list)

    @property
    def n_owned(self) -> int:
        return self.node_hi - self.node_lo

    @property
    def n_halo(self) -> int:
        return int(self.halo_ids.numel())

    @property
    def n_local(self) -> int:
        return self.n_owned + self.n_halo

    def to(self, device) -> "GraphShard":
        return GraphShard(self.rank, self.world, self.node_lo, self.node_hi, self.halo_ids.to(device),
                          self.edge_ids.to(device), self.edge_index.to(device), self.send_idx.to(device),
                          list(self.send_counts), list(self.recv_counts))


def node_ranges(n_nodes: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, near-equal ranges: rank p owns [lo_p, hi_p)."""
    base, rem = divmod(n_nodes, world)
    out, lo = [], 0
    for p in range(world):
        hi = lo + base + (1 if p < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def owner_of(ids: Tensor, n_nodes: int, world: int) -> Tensor:
    base, rem = divmod(n_nodes, world)
    cut = rem * (base + 1)  # the first `rem` ranks own base + 1 nodes
    lo_part = torch.div(ids, base + 1, rounding_mode="floor")
    hi_part = rem + torch.div(ids - cut, max(base, 1), rounding_mode="floor")
    return torch.where(ids < cut, lo_part, hi_part)


def partition_graph(edge_index: Tensor, n_nodes: int, world: int, rank: int) -> GraphShard:
    """Shard of ``rank``.  Every rank calls this on the same (global) ``edge_index``: the halo
    of every pair of ranks is derived from it deterministically, so no set-up communication."""
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise TypeError("edge_index must be an int64 tensor of shape [2, E]")
    src, dst = edge_index[0], edge_index[1]
    ranges = node_ranges(n_nodes, world)
    lo, hi = ranges[rank]
    dst_owner = owner_of(dst, n_nodes, world)
    src_owner = owner_of(src, n_nodes, world)
    mine = dst_owner == rank
    edge_ids = torch.nonzero(mine).flatten()
    s, d = src[edge_ids], dst[edge_ids]
    remote = (s < lo) | (s >= hi)
    halo_ids = torch.unique(s[remote])  # sorted
    local_src = torch.where(remote, (hi - lo) + torch.searchsorted(halo_ids, s), s - lo)
    recv_counts = [0] * world
    if halo_ids.numel():
        ho = owner_of(halo_ids, n_nodes, world)
        recv_counts = torch.bincount(ho, minlength=world).tolist()
    # what the others need from me: unique sources I own among the edges THEY own
    send_idx, send_counts = [], []
    theirs_from_me = (src_owner == rank) & ~mine
    for q in range(world):
        if q == rank:
            send_counts.append(0)
            continue
        need = torch.unique(src[theirs_from_me & (dst_owner == q)])
        send_idx.append((need - lo).to(torch.int32))
        send_counts.append(int(need.numel()))
    send = torch.cat(send_idx) if send_idx else torch.zeros(0, dtype=torch.int32)
    return GraphShard(rank, world, lo, hi, halo_ids, edge_ids, torch.stack([local_src, d - lo]), send,
                      send_counts, recv_counts)


class HaloExchange:
    """Per-layer exchange of halo rows: ``extend(table [n_owned, w]) -> [n_owned + n_halo, w]``.

    One all-to-all-v as a single batch of point-to-point operations (NCCL fuses the batch into
    one grouped launch over NVLink; no staging through the host)."""

    def __init__(self, shard: GraphShard, group=None):
        self.shard = shard
        self.group = group
        self.bytes_sent = 0

    def buffer(self, width: int, like: Tensor) -> Tensor:
        """Table for ``n_owned + n_halo`` rows; kernels write the owned rows in place."""
        return torch.empty((self.shard.n_local, width), dtype=like.dtype, device=like.device)

    def extend(self, table: Tensor, buf: Tensor | None = None) -> Tensor:
        return self.finish(self.start(table, buf))

    def start(self, table: Tensor, buf: Tensor | None = None):
        """Issues the exchange (asynchronously where the backend allows) and returns a handle for
        ``finish``: the caller launches whatever does not need the halo rows in between -- the other
        pre-projection of the layer runs under the transfer."""
        sh = self.shard
        if table.dim() != 2 or table.size(0) not in (sh.n_owned, sh.n_local):
            raise ValueError(f"expected a table of {sh.n_owned} owned rows, got {tuple(table.shape)}")
        w = table.size(1)
        if table.size(0) == sh.n_local:   # already the extended buffer: owned rows are in place
            buf = table
        else:
            if buf is None:
                buf = self.buffer(w, table)
            buf[:sh.n_owned].copy_(table)
        if sh.world == 1:
            return buf, None
        send = self._pack(buf, sh.send_idx, sh.n_owned)
        self.bytes_sent += send.numel() * send.element_size()
        if buf.is_cuda:
            # one all-to-all-v: NCCL runs it on its own stream, the caller's stream goes on
            work = dist.all_to_all_single(buf[sh.n_owned:], send, output_split_sizes=list(sh.recv_counts),
                                          input_split_sizes=list(sh.send_counts), group=self.group, async_op=True)
            return buf, (work, send)
        ops_, so, ro = [], 0, sh.n_owned  # gloo (CPU tests): the same exchange as point-to-point operations
        for q in range(sh.world):
            if sh.recv_counts[q]:
                ops_.append(dist.P2POp(dist.irecv, buf[ro:ro + sh.recv_counts[q]], self._peer(q), self.group))
                ro += sh.recv_counts[q]
            if sh.send_counts[q]:
                ops_.append(dist.P2POp(dist.isend, send[so:so + sh.send_counts[q]], self._peer(q), self.group))
                so += sh.send_counts[q]
        return buf, (dist.batch_isend_irecv(ops_) if ops_ else [], send)

    @staticmethod
    def finish(handle) -> Tensor:
        buf, pending = handle
        if pending is not None:
            work, _send = pending
            for req in (work if isinstance(work, list) else [work]):
                req.wait()  # NCCL: the current stream waits for the transfer, the host does not
        return buf

    def reverse(self, grad_ext: Tensor) -> Tensor:
        """Adjoint of ``extend``: ``grad_ext [n_owned + n_halo, w] -> grad [n_owned, w]``.  The halo rows'
        gradients travel back to their owners (the reverse all-to-all-v) and are added onto the
        rows they were copied from; a row sent to several ranks collects all of them."""
        sh = self.shard
        g = grad_ext.contiguous()
        out = g[:sh.n_owned].clone()
        if sh.world == 1 or (sh.n_halo == 0 and sum(sh.send_counts) == 0):
            return out
        back = g[sh.n_owned:]
        recv = torch.empty((sum(sh.send_counts), g.size(1)), dtype=g.dtype, device=g.device)
        if g.is_cuda:
            dist.all_to_all_single(recv, back, output_split_sizes=list(sh.send_counts),
                                   input_split_sizes=list(sh.recv_counts), group=self.group)
        else:  # gloo (CPU tests)
            ops_, so, ro = [], 0, 0
            for q in range(sh.world):
                if sh.send_counts[q]:
                    ops_.append(dist.P2POp(dist.irecv, recv[ro:ro + sh.send_counts[q]], self._peer(q), self.group))
                    ro += sh.send_counts[q]
                if sh.recv_counts[q]:
                    ops_.append(dist.P2POp(dist.isend, back[so:so + sh.recv_counts[q]].contiguous(), self._peer(q), self.group))
                    so += sh.recv_counts[q]
            for req in (dist.batch_isend_irecv(ops_) if ops_ else []):
                req.wait()
        self.bytes_sent += back.numel() * back.element_size()
        if recv.numel():
            if g.is_cuda:
                from . import ops
                ops.rows_scatter_add(recv, sh.send_idx, out)
            else:
                out.index_add_(0, sh.send_idx.long(), recv)
        return out

    def _peer(self, q: int) -> int:
        return q if self.group is None else dist.get_global_rank(self.group, q)

    @staticmethod
    def _pack(buf: Tensor, idx: Tensor, n_owned: int) -> Tensor:
        if buf.is_cuda:
            from . import ops
            return ops.rows_gather(buf, idx)
        return buf.index_select(0, idx.long())  # CPU tensors: host-side tests of the plumbing (gloo)


class _HaloExtend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table: Tensor, halo: "HaloExchange") -> Tensor:
        ctx.halo = halo
        return halo.extend(table.detach())

    @staticmethod
    def backward(ctx, grad_ext: Tensor):
        return ctx.halo.reverse(grad_ext), None


def halo_extend(table: Tensor, halo: HaloExchange) -> Tensor:
    """Differentiable ``halo.extend(table)``: the backward is the reverse exchange (``HaloExchange.reverse``).
    Training on a node-partitioned graph exchanges the RAW node rows of a layer through this (the fused
    kernel's backward then treats the extended table like any gathered table)."""
    return _HaloExtend.apply(table, halo)


def allreduce_gradients(module: torch.nn.Module, group=None) -> None:
    """Sums the parameter gradients over the ranks of a node-partitioned graph (every rank holds all the
    weights and the gradient of ITS edges and nodes): call between ``backward()`` and the optimiser
    step.  Losses must be normalised with GLOBAL counts (e.g. BCE summed locally / the global edge count)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()

