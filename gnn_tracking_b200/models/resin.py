"""Residual stacks of interaction networks (reference models/resin.py:45-295) over
one shared graph plan.  The residual combination ``sqrt(a) x + sqrt(1-a) dx``
(resin.py:17-42) and the ReLU in front of every layer but the first
(resin.py:104-105) are fused into the layer kernels (epilogue / on-load
activation), so a layer is exactly two launches and no elementwise pass."""
from __future__ import annotations

import math
import os
from abc import ABC, abstractmethod
from itertools import pairwise

import torch
from torch import Tensor, nn

from .. import ops
from .._hparams import HyperparametersMixin
from ..plan import GraphPlan, get_plan
from .interaction_network import InteractionNetwork
from .mlp import autocast_bf16


def _res_coeffs(alpha: float) -> tuple[float, float] | None:
    """None when the residual is skipped (alpha ~ 0, resin.py:38-39)."""
    if math.isclose(alpha, 0.0):
        return None
    return math.sqrt(alpha), math.sqrt(1.0 - alpha)


def mark_sorted_edges(t: Tensor) -> Tensor:
    """Tags an edge tensor whose rows follow the plan's destination-sorted edge order (row i = edge
    ``plan.perm[i]``).  A plain attribute on the tensor object: it does not survive any torch op, so
    only the very tensors produced inside ``forward_planned(sorted_edges=True)`` carry it."""
    t._gtb_sorted_edges = True
    return t


def has_sorted_edges(t: Tensor) -> bool:
    return getattr(t, "_gtb_sorted_edges", False)


class ResidualNetwork(ABC, nn.Module):
    def __init__(self, layers: list[nn.Module], *, alpha: float = 0.5, collect_hidden_edge_embeds: bool = False):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self._alpha = alpha
        self._collect_hidden_edge_embeds = collect_hidden_edge_embeds

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> tuple[Tensor, Tensor, list[Tensor] | None]:
        return self._forward(x, get_plan(edge_index, x.size(0)), edge_attr)

    def forward_planned(self, x: Tensor, plan: GraphPlan, edge_attr: Tensor, halo=None, sorted_edges: bool = False,
                        final_projection=None):
        """``sorted_edges``: ``edge_attr`` is given in the plan's destination-sorted edge order and every
        edge tensor of the stack EXCEPT the final one stays in that order (contiguous tiles instead of a
        gather / scatter through ``perm`` per layer).  The final ``edge_attr`` is in the caller's order
        as always; the entries of the returned ``edge_attrs`` list answer ``has_sorted_edges`` where
        they are in sorted order.

        ``final_projection``: ``(pair of packed projections, position of the SOURCE-gathered block in the
        pair)`` -- ``mlp.projection_packs`` of the consumer of the final node embedding: the W head.  When the
        stack runs its fused node launches the last one computes them too and ``self.final_tables`` holds
        the two tables afterwards (else None)."""
        self._halo = halo
        self._sorted = sorted_edges and len(self.layers) > 0
        self._calls_left = self._n_layer_calls()
        self.final_tables = None
        self._fused = self._fused_state(x, final_projection) if self.fused_ok(x, edge_attr, halo) else None
        try:
            if self._sorted:
                edge_attr = mark_sorted_edges(edge_attr)
            out = self._forward(x, plan, edge_attr)
            if self._fused is not None:  # every node launch handed the aggregate back zeroed: keep it for the next graph
                self.__dict__["_aggr_buf"] = self._fused["aggr"]
            return out
        finally:
            self._halo = None
            self._sorted = False
            self._fused = None

    _halo = None
    _sorted = False
    _calls_left = 0
    _fused = None
    final_tables = None

    def _n_layer_calls(self) -> int:
        return len(self.layers)

    def _call_sequence(self) -> list[int]:
        """Layer index of every ``_layer`` call of ``_forward``, in order."""
        return list(range(len(self.layers)))

    def fused_ok(self, x: Tensor, edge_attr: Tensor, halo=None) -> bool:
        """Does the two-launches-per-layer path (``InteractionNetwork.forward_fused``) apply?  fp32 without
        autograd, every layer 64 / 64 / 64 on the tensor-core tiles, at least two edges per (owned) node
        (the condition under which the node blocks are pre-projected at all)."""
        if (len(self.layers) == 0 or os.environ.get("GTB_NO_NODE_WS") or autocast_bf16()
                or 2 * x.size(0) > edge_attr.size(0) or x.size(0) == 0 or ops.default_impl() == ops.IMPL_FFMA):
            return False
        if torch.is_grad_enabled() and (x.requires_grad or edge_attr.requires_grad
                                        or any(p.requires_grad for p in self.parameters())):
            return False
        return all(isinstance(l, InteractionNetwork) and l.fused_wide() for l in self.layers)

    def _fused_state(self, x: Tensor, final_projection):
        seq = self._call_sequence()
        projs = {i: self.layers[i].input_projection() for i in set(seq)}
        if any(v is None for v in projs.values()):
            return None
        # the zero-filled aggregate of the previous forward is reused when it fits (popped: an exception in the
        # middle of a stack leaves it dirty and it is then simply dropped)
        aggr = self.__dict__.pop("_aggr_buf", None)
        if aggr is None or aggr.size(0) != x.size(0) or aggr.device != x.device:
            aggr = torch.zeros((x.size(0), 64), dtype=torch.float32, device=x.device)
        final, side = final_projection if final_projection is not None else (None, 1)
        return {"seq": seq, "k": 0, "projs": projs, "final": final, "final_src_side": side, "tables": None, "x": None,
                "aggr": aggr}

    def _layer(self, i: int, x: Tensor, plan: GraphPlan, e: Tensor, *, first: bool, residue: Tensor | None):
        """IN layer i on (act(x), act(e)) with the residual onto the un-activated
        ``residue`` fused in; act = identity for the very first layer, else ReLU."""
        co = _res_coeffs(self._alpha) if residue is not None else None
        kw = {} if co is None else dict(res=residue, res_a=co[0], res_b=co[1])
        self._calls_left -= 1
        out_sorted = self._sorted and self._calls_left > 0  # the stack's final edge tensor: caller's order
        torch.cuda.nvtx.range_push(f"gtb.in_layer.{i}")
        try:
            xo, eo = self._layer_call(i, x, plan, e, first, out_sorted, kw)
        finally:
            torch.cuda.nvtx.range_pop()
        return xo, (mark_sorted_edges(eo) if out_sorted else eo)

    def _layer_call(self, i, x, plan, e, first, out_sorted, kw):
        f = self._fused
        if f is None:
            return self.layers[i].forward_planned(x, plan, e, relu_x=not first, relu_e=not first, halo=self._halo,
                                                    e_sorted=has_sorted_edges(e), out_sorted=out_sorted, **kw)
        k = f["k"]
        assert f["seq"][k] == i
        last = k + 1 == len(f["seq"])
        nxt = f["final"] if last else f["projs"][f["seq"][k + 1]]
        # the tables computed by the previous call belong to ITS output: any other input projects for itself
        tables = f["tables"] if f["x"] is x else None
        xo, eo, nt = self.layers[i].forward_fused(
            x, plan, e, relu_x=not first, relu_e=not first, res=kw.get("res"), res_a=kw.get("res_a", 0.0),
            res_b=kw.get("res_b", 1.0), e_sorted=has_sorted_edges(e), out_sorted=out_sorted, tables=tables,
            aggr=f["aggr"], nxt=nxt, nxt_relu=not last, nxt_src_side=f["final_src_side"] if last else 1, halo=self._halo)
        f["k"], f["tables"], f["x"] = k + 1, nt, xo
        if last:
            self.final_tables = nt
        return xo, eo

    @abstractmethod
    def _forward(self, x: Tensor, plan: GraphPlan, edge_attr: Tensor):
        ...


class Skip1ResidualNetwork(ResidualNetwork):
    """Every layer has a residual connection to its input (resin.py:99-114)."""

    def _forward(self, x, plan, edge_attr):
        edge_attrs = [edge_attr] if self._collect_hidden_edge_embeds else None
        for i in range(len(self.layers)):
            x, edge_attr = self._layer(i, x, plan, edge_attr, first=i == 0, residue=x)
            if edge_attrs is not None:
                edge_attrs.append(edge_attr)
        return x, edge_attr, edge_attrs


class Skip2ResidualNetwork(ResidualNetwork):
    """Blocks of two layers with a residual connection around each block
    (resin.py:153-175).  The reference iterates ``pairwise(range(L))`` -- overlapping
    pairs -- which is reproduced literally."""

    def __init__(self, layers: list[nn.Module], *, node_dim: int, edge_dim: int, add_bn: bool = False, **kwargs):
        if len(layers) % 2 != 0:
            raise ValueError("Only even number of layers allowed at the moment")
        super().__init__(layers=layers, **kwargs)
        # the reference's module tree (resin.py:141-151): BatchNorm1d per layer input with ``add_bn``, else
        # parameter-free placeholders.  The normalisation itself is the library's (torch.nn.BatchNorm1d: a
        # column-wise affine map of an N x D / E x D table in front of a layer, batch statistics in training);
        # it is row-permutation invariant, so it applies to destination-sorted edge tables unchanged.
        self._add_bn = bool(add_bn)
        self._node_batch_norms = nn.ModuleList([nn.BatchNorm1d(node_dim) if add_bn else nn.Identity() for _ in layers])
        self._edge_batch_norms = nn.ModuleList([nn.BatchNorm1d(edge_dim) if add_bn else nn.Identity() for _ in layers])

    def _n_layer_calls(self) -> int:
        return 2 * max(len(self.layers) - 1, 0)

    def _call_sequence(self) -> list[int]:
        return [i for pair in pairwise(range(len(self.layers))) for i in pair]

    def fused_ok(self, x, edge_attr, halo=None) -> bool:
        return not self._add_bn and super().fused_ok(x, edge_attr, halo)  # the node launch hands x on un-normalised

    def _bn(self, norms, i: int, t: Tensor) -> Tensor:
        if not self._add_bn:
            return t
        out = norms[i](t)
        return mark_sorted_edges(out) if has_sorted_edges(t) else out

    def _forward(self, x, plan, edge_attr):
        edge_attrs = [edge_attr] if self._collect_hidden_edge_embeds else None
        for i0, i1 in pairwise(range(len(self.layers))):
            # resin.py:157-168: act(bn(.)) in front of both layers; the residue is the un-normalised x
            hx, he = self._layer(i0, self._bn(self._node_batch_norms, i0, x), plan, self._bn(self._edge_batch_norms, i0, edge_attr),
                                 first=i0 == 0, residue=None)
            x, edge_attr = self._layer(i1, self._bn(self._node_batch_norms, i1, hx), plan, self._bn(self._edge_batch_norms, i1, he),
                                       first=False, residue=x)
            if edge_attrs is not None:
                edge_attrs.append(edge_attr)
        return x, edge_attr, edge_attrs


class SkipTopResidualNetwork(ResidualNetwork):
    """Residual connections to one fixed early layer output (resin.py:197-216)."""

    def __init__(self, layers: list[nn.Module], connect_to: int = 1, **kwargs):
        assert connect_to <= len(layers)
        super().__init__(layers=layers, **kwargs)
        self._residual_layer = connect_to

    def _forward(self, x, plan, edge_attr):
        edge_attrs = [edge_attr] if self._collect_hidden_edge_embeds else None
        x_residue = None
        for i in range(len(self.layers)):
            if i == self._residual_layer:
                x_residue = x
            x, edge_attr = self._layer(i, x, plan, edge_attr, first=i == 0, residue=x_residue)
            if edge_attrs is not None:
                edge_attrs.append(edge_attr)
        return x, edge_attr, edge_attrs


RESIDUAL_NETWORKS_BY_NAME = {
    "skip1": Skip1ResidualNetwork,
    "skip2": Skip2ResidualNetwork,
    "skip_top": SkipTopResidualNetwork,
}


class ResIN(nn.Module, HyperparametersMixin):
    def __init__(self, *, node_dim: int, edge_dim: int, object_hidden_dim=40, relational_hidden_dim=40,
                 alpha: float = 0.5, n_layers=1, residual_type: str = "skip1",
                 residual_kwargs: dict | None = None):
        """``n_layers`` identical interaction networks inside the residual network named
        by ``residual_type`` (reference resin.py:226-281; parameters under
        ``network.layers.{i}.*``)."""
        super().__init__()
        self.save_hyperparameters()
        residual_kwargs = dict(residual_kwargs or {})
        layers = [InteractionNetwork(node_indim=node_dim, edge_indim=edge_dim, node_outdim=node_dim,
                                     edge_outdim=edge_dim, node_hidden_dim=object_hidden_dim,
                                     edge_hidden_dim=relational_hidden_dim) for _ in range(n_layers)]
        if residual_type == "skip2":
            residual_kwargs["node_dim"] = node_dim
            residual_kwargs["edge_dim"] = edge_dim
        self.network = RESIDUAL_NETWORKS_BY_NAME[residual_type](layers, alpha=alpha, **residual_kwargs)
        self.node_dim = node_dim
        self.edge_dim = edge_dim
        self._residual_type = residual_type

    @property
    def concat_edge_embeddings_length(self) -> int:
        """Width of the concatenated per-layer edge embeddings (resin.py:283-290)."""
        n = len(self.network.layers)
        return self.edge_dim * ((n // 2 if self._residual_type == "skip2" else n) + 1)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor):
        return self.network.forward(x, edge_index, edge_attr)

    def forward_planned(self, x: Tensor, plan: GraphPlan, edge_attr: Tensor, halo=None, sorted_edges: bool = False,
                        final_projection=None):
        return self.network.forward_planned(x, plan, edge_attr, halo=halo, sorted_edges=sorted_edges,
                                            final_projection=final_projection)
