"""Interaction-Network layer behind the reference's module interface
(reference models/interaction_network.py:12-103): same constructor arguments,
``.hparams``, parameter names (``relational_model.layers.*``,
``object_model.layers.*``) and ``forward(x, edge_index, edge_attr) -> (x_tilde,
e_tilde)`` in the caller's edge order -- computed by two fused kernels over a
destination-sorted plan instead of PyG's gather / cat / addmm / scatter_add chain."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from .._hparams import HyperparametersMixin
from ..ops import Block
from ..plan import GraphPlan, get_plan
from ..utils.asserts import assert_feat_dim
from .mlp import MLP, autocast_bf16, projection_packs


def _project_pair(x: Tensor, fused: dict | None, packs, proj_relu: bool, src_side: int, halo, relu_x: bool = False):
    """One ``ops.in_node_fused`` launch -- projection only (``fused`` None: returns the pair of tables of
    ``act(x)``) or object model + projections (returns ``(x_out, pair)``) -- with the source-side table
    of the pair written into the halo exchange's extended table and completed by the exchange."""
    outs = [None, None]
    if halo is not None:
        outs[src_side] = halo.buffer(64, x)
    xo, pa, pb = ops.in_node_fused(x, relu_x, proj=packs, proj_relu=proj_relu, out_pa=outs[0], out_pb=outs[1],
                                   **(fused or {}))
    pair = [pa, pb]
    if halo is not None:
        pair[src_side] = halo.finish(halo.start(pair[src_side]))
    return tuple(pair) if fused is None else (xo, tuple(pair))


class InteractionNetwork(nn.Module, HyperparametersMixin):
    def __init__(self, *, node_indim: int, edge_indim: int, node_outdim: int = 3, edge_outdim: int = 4,
                 node_hidden_dim: int = 40, edge_hidden_dim: int = 40, aggr: str = "add"):
        super().__init__()
        self.save_hyperparameters()
        if aggr != "add":
            raise ValueError("only aggr='add' (the reference default, used by every model) is implemented")
        self.relational_model = MLP(2 * node_indim + edge_indim, edge_outdim, edge_hidden_dim)
        self.object_model = MLP(node_indim + edge_outdim, node_outdim, node_hidden_dim)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> tuple[Tensor, Tensor]:
        assert_feat_dim(x, self.hparams.node_indim)
        assert_feat_dim(edge_attr, self.hparams.edge_indim)
        plan = get_plan(edge_index, x.size(0))
        return self.forward_planned(x, plan, edge_attr)

    def forward_planned(self, x: Tensor, plan: GraphPlan, edge_attr: Tensor, *, relu_x: bool = False,
                        relu_e: bool = False, res: Tensor | None = None, res_a: float = 0.0,
                        res_b: float = 1.0, halo=None, e_sorted: bool = False,
                        out_sorted: bool = False) -> tuple[Tensor, Tensor]:
        """One layer on a planned graph.  ``relu_*`` apply the activation on load (layers
        > 0 of a residual stack see relu(x), reference models/resin.py:104-105); ``res``
        fuses ``sqconvex_combination`` (resin.py:17-42) into the node kernel.  ``halo`` (a
        ``partition.HaloExchange``): ``x`` holds the owned nodes of a node-partitioned graph and
        ``plan`` its local edges; the source rows other ranks own are exchanged once per layer.
        ``e_sorted`` / ``out_sorted``: ``edge_attr`` is given / ``e_tilde`` is returned in the plan's
        destination-sorted edge order instead of the caller's (row i = edge ``plan.perm[i]``): inside a
        stack the edge features then stream as contiguous tiles, no gather or scatter through ``perm``
        (the module-level ``forward`` never exposes that order)."""
        dev = ops.require_cuda(x, edge_attr)
        n, e = x.size(0), edge_attr.size(0)
        if autocast_bf16() and halo is None and self._bf16_wide():
            return self._forward_bf16(x, plan, edge_attr, relu_x=relu_x, relu_e=relu_e, res=res, res_a=res_a,
                                      res_b=res_b, e_sorted=e_sorted, out_sorted=out_sorted)
        e_out = self.hparams.edge_outdim
        # zeroed: isolated nodes keep 0 (SumAggregation), partial runs are added atomically
        ext = None if halo is None else halo.extend
        # message(): cat[x_i (target), x_j (source), edge_attr]  (interaction_network.py:75-89); the
        # aggregate starts from zero: isolated nodes keep 0 (SumAggregation)
        e_tilde, aggr = self.relational_model.forward_blocks(
            [Block(x, plan.dst_sorted, relu_x, sorted_index=True), Block(x, plan.src_sorted, relu_x, extend=ext),
             Block(edge_attr, None, relu_e) if e_sorted else Block(edge_attr, plan.perm, relu_e, unique_index=True)],
            e, out_index=None if out_sorted else plan.perm, aggr_rows=n, seg_id=plan.dst_sorted, rowptr=plan.rowptr)
        # update(): cat[x, aggr]  (interaction_network.py:92-103)
        x_tilde = self.object_model.forward_blocks([Block(x, None, relu_x), Block(aggr)], n,
                                                   res=res, res_a=res_a, res_b=res_b)
        return x_tilde, e_tilde

    # ------------------------------------------------------------------ fused node side (64-wide layers, no grad)
    def fused_wide(self) -> bool:
        """True when both models have the 64 / 64 / 64 shape ``forward_fused`` covers."""
        hp = self.hparams
        return (hp.node_indim == hp.edge_indim == hp.node_outdim == hp.edge_outdim == hp.node_hidden_dim
                == hp.edge_hidden_dim == 64 and len(self.relational_model.linears) == 3
                and len(self.object_model.linears) == 3)

    def input_projection(self):
        """The packed projections of this layer's two gathered node blocks (``ops.in_node_fused`` of the
        PREVIOUS layer computes them), or None."""
        return projection_packs(self.relational_model, (64, 64, 64))

    def forward_fused(self, x: Tensor, plan: GraphPlan, edge_attr: Tensor, *, relu_x: bool, relu_e: bool,
                      res: Tensor | None, res_a: float, res_b: float, e_sorted: bool, out_sorted: bool,
                      tables: tuple[Tensor, Tensor] | None, aggr: Tensor, nxt=None, nxt_relu: bool = True,
                      nxt_src_side: int = 1, halo=None):
        """The layer as two launches (no-grad fp32, ``fused_wide()`` shapes): the edge kernel on the
        pre-projected node tables ``tables`` = (P_i, P_j) it is handed, and ONE node launch
        (``ops.in_node_fused``) that runs the object model with the residual, hands ``aggr`` back zeroed
        and computes the tables of the next consumer ``nxt`` (a pair of packed projections: the next
        layer's ``input_projection()`` or the W head's).  Returns ``(x_tilde, e_tilde, next_tables)``.

        ``halo`` (node-partitioned graph): the table gathered by SOURCE ids -- position ``nxt_src_side`` of
        the next consumer's pair, position 1 of this layer's -- is written into the extended table of the
        exchange and completed with the other ranks' rows (one all-to-all-v) before it is handed on."""
        n, e = x.size(0), edge_attr.size(0)
        if tables is None:  # first layer of a stack: projection-only launch
            tables = _project_pair(x, None, self.input_projection(), relu_x, 1, halo)
        e_tilde = self.relational_model.forward_blocks(
            [Block(x, plan.dst_sorted, relu_x, sorted_index=True),
             Block(x, plan.src_sorted, relu_x, extend=None if halo is None else halo.extend),
             Block(edge_attr, None, relu_e) if e_sorted else Block(edge_attr, plan.perm, relu_e, unique_index=True)],
            e, out_index=None if out_sorted else plan.perm, aggr=aggr, seg_id=plan.dst_sorted, rowptr=plan.rowptr,
            tables={0: tables[0], 1: tables[1]})
        obj = self.object_model
        packed_obj = obj._cache.get(obj.linears, (64, 64), (False, False))[0][0]
        fused = dict(aggr=aggr, zero_aggr=True, packed_obj=packed_obj, res=res, res_a=res_a, res_b=res_b)
        if nxt is None:
            return ops.in_node_fused(x, relu_x, **fused)[0], e_tilde, None
        x_tilde, nt = _project_pair(x, fused, nxt, nxt_relu, nxt_src_side, halo, relu_x=relu_x)
        return x_tilde, e_tilde, nt

    # ------------------------------------------------------------------ bf16 (torch.autocast) path
    def _bf16_wide(self) -> bool:
        hp = self.hparams
        return (hp.node_indim == hp.edge_indim == hp.edge_outdim == hp.edge_hidden_dim == 128)

    def _forward_bf16(self, x, plan, edge_attr, *, relu_x, relu_e, res, res_a, res_b, e_sorted, out_sorted):
        """The layer under ``torch.autocast(bfloat16)`` at width 128 (BASELINE config 3): the edge side --
        gather, relational MLP, per-destination sum -- is ONE native bf16 launch (``gtb_in_edge_forward_bf16``:
        tcgen05 kind::f16, fp32 accumulation, Linear outputs rounded to bf16 as autocast does); the node
        side (the two N x 128 x 128 pre-projections and the object model) are plain library GEMMs on
        bf16 operands."""
        if torch.is_grad_enabled() and (x.requires_grad or edge_attr.requires_grad
                                        or any(p.requires_grad for p in self.relational_model.parameters())):
            raise NotImplementedError("the bf16 edge kernel is forward-only: wrap inference in torch.no_grad()")
        bf = torch.bfloat16
        rel = self.relational_model.linears
        w0 = rel[0].weight
        cache = self.__dict__.setdefault("_bf16_cache", {})
        key = tuple((p.data_ptr(), p._version) for lin in rel for p in lin.parameters())
        if cache.get("key") != key:
            cache["packed"] = ops.pack_in_edge_bf16([w0[:, 256:].contiguous(), rel[1].weight, rel[2].weight],
                                                    [rel[0].bias, rel[1].bias, rel[2].bias])
            cache["w_i"] = w0[:, :128].detach().to(bf).contiguous()
            cache["w_j"] = w0[:, 128:256].detach().to(bf).contiguous()
            cache["key"] = key
        xa = (torch.relu(x) if relu_x else x).to(bf)
        p_i = nn.functional.linear(xa, cache["w_i"])
        p_j = nn.functional.linear(xa, cache["w_j"])
        ea = edge_attr.to(bf)
        if not ea.is_contiguous():
            ea = ea.contiguous()
        e_tilde, aggr = ops.in_edge_bf16(ea, p_i, p_j, plan.src_sorted, plan.dst_sorted, cache["packed"], x.size(0),
                                         e_index=None if e_sorted else plan.perm,
                                         out_index=None if out_sorted else plan.perm, relu_e=relu_e)
        # update(): cat[x, aggr] promotes to fp32, the Linear casts back to bf16 (interaction_network.py:92-103)
        h = torch.cat([xa.float(), aggr], dim=1)
        for m in self.object_model.layers:
            h = m(h)
        if res is not None:
            h = res_a * res + res_b * h
        return h, e_tilde

