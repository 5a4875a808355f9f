"""Edge classifier of the Graph-TCN behind the reference interface (reference
models/edge_classifier.py:15-121): encoders -> ResIN -> W head, each a fused
kernel over the shared destination-sorted plan."""
from __future__ import annotations

import math

import torch
from torch import Tensor, nn

from .. import ops
from .._hparams import HyperparametersMixin
from ..ops import ACT_RELU, ACT_SIGMOID_AFFINE, Block
from ..plan import GraphPlan, get_plan
from ..utils.asserts import assert_feat_dim
from .mlp import MLP, projection_packs
from .resin import ResIN, has_sorted_edges


class ECForGraphTCN(nn.Module, HyperparametersMixin):
    def __init__(self, *, node_indim: int, edge_indim: int, interaction_node_dim: int = 5,
                 interaction_edge_dim: int = 4, hidden_dim: int | float | None = None, L_ec: int = 3,
                 alpha: float = 0.5, residual_type="skip1", use_intermediate_edge_embeddings: bool = True,
                 use_node_embedding: bool = True, residual_kwargs: dict | None = None):
        """Same arguments, ``.hparams`` and parameter names as the reference
        (``ec_node_encoder``, ``ec_edge_encoder``, ``ec_resin.network.layers.*``, ``W``)."""
        super().__init__()
        self.save_hyperparameters()
        residual_kwargs = dict(residual_kwargs or {})
        residual_kwargs["collect_hidden_edge_embeds"] = use_intermediate_edge_embeddings
        self.relu = nn.ReLU()
        self.ec_node_encoder = MLP(node_indim, interaction_node_dim, hidden_dim=hidden_dim, L=2, bias=False)
        self.ec_edge_encoder = MLP(edge_indim, interaction_edge_dim, hidden_dim=hidden_dim, L=2, bias=False)
        self.ec_resin = ResIN(node_dim=interaction_node_dim, edge_dim=interaction_edge_dim,
                              object_hidden_dim=hidden_dim, relational_hidden_dim=hidden_dim, alpha=alpha,
                              n_layers=L_ec, residual_type=residual_type, residual_kwargs=residual_kwargs)
        w_in = interaction_edge_dim
        if use_intermediate_edge_embeddings:
            w_in = self.ec_resin.concat_edge_embeddings_length
        if use_node_embedding:
            w_in += 2 * interaction_node_dim
        self.W = MLP(input_size=w_in, output_size=1, hidden_dim=hidden_dim, L=3)
        self.latent_dim = (interaction_node_dim, interaction_edge_dim)

    def forward(self, data) -> dict[str, Tensor]:
        x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        return self.forward_tensors(x, edge_index, edge_attr)

    def forward_tensors(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor,
                        plan: GraphPlan | None = None, halo=None) -> dict[str, Tensor]:
        """``halo`` (``partition.HaloExchange``): ``x`` / ``edge_attr`` are the owned nodes / edges of
        a node-partitioned graph and ``edge_index`` its local numbering (``GraphShard.edge_index``);
        the outputs cover the owned edges and nodes."""
        assert_feat_dim(x, self.hparams.node_indim)
        assert_feat_dim(edge_attr, self.hparams.edge_indim)
        ops.require_cuda(x, edge_index, edge_attr)
        if plan is None:
            torch.cuda.nvtx.range_push("gtb.plan")
            plan = get_plan(edge_index, x.size(0) if halo is None else halo.shard.n_local)
            torch.cuda.nvtx.range_pop()
        n, e = x.size(0), edge_attr.size(0)
        nvtx = torch.cuda.nvtx
        # encoders + the ReLU behind them (edge_classifier.py:102-103)
        nvtx.range_push("gtb.ec.encoders")
        h = self.ec_node_encoder.forward_blocks([Block(x)], n, final_act=ACT_RELU)
        # The edge features live in the plan's destination-sorted order inside the model (the encoder
        # gathers its 4 input columns through ``perm``): every layer then streams contiguous tiles.  Only
        # the final embedding -- the one the caller sees -- is written back in the caller's order.
        sorted_edges = len(self.ec_resin.network.layers) > 0
        if self.ec_edge_encoder.k4_ok(edge_attr):  # 4 -> 64 -> 64: the dedicated one-launch encoder
            ea = self.ec_edge_encoder.forward_k4(edge_attr, plan.perm if sorted_edges else None, e, final_relu=True)
        else:
            ea = self.ec_edge_encoder.forward_blocks(
                [Block(edge_attr, plan.perm, unique_index=True) if sorted_edges else Block(edge_attr)], e, final_act=ACT_RELU)
        nvtx.range_pop()
        nvtx.range_push("gtb.ec.resin")
        # the last node launch of the stack also multiplies the final node embedding by the head's two node
        # column blocks (the products the head gathers): no separate projection launches
        hp = self.hparams
        n_eb = (self.ec_resin.concat_edge_embeddings_length // hp.interaction_edge_dim
                if hp.use_intermediate_edge_embeddings else 1)
        final_proj = None
        if hp.use_node_embedding and self.ec_resin.network.fused_ok(h, ea, halo):
            packs = projection_packs(self.W, [hp.interaction_node_dim] * 2 + [hp.interaction_edge_dim] * n_eb)
            final_proj = None if packs is None else (packs, 0)  # the head's blocks: [h[src], h[dst], e_0 ..]
        h, ea, eas = self.ec_resin.forward_planned(h, plan, ea, halo=halo, sorted_edges=sorted_edges,
                                                   final_projection=final_proj)
        head_tables = self.ec_resin.network.final_tables
        nvtx.range_pop()
        nvtx.range_push("gtb.ec.w_head")
        # W head over cat[h[src], h[dst], e_0 .. e_L] (edge_classifier.py:108-117), walked in
        # dst-sorted order (h[dst] rows repeat) and written back in the caller's edge order
        blocks = []
        if self.hparams.use_node_embedding:
            blocks += [Block(h, plan.src_sorted, extend=None if halo is None else halo.extend),
                       Block(h, plan.dst_sorted, sorted_index=True)]
        blocks += [Block(t, None) if has_sorted_edges(t) else Block(t, plan.perm, unique_index=True)
                   for t in (eas if self.hparams.use_intermediate_edge_embeddings else [ea])]
        kw = {} if head_tables is None else {"tables": {0: head_tables[0], 1: head_tables[1]}}
        w = self.W.forward_blocks(blocks, e, final_act=ACT_SIGMOID_AFFINE, act_eps=0.001, out_index=plan.perm, **kw)
        nvtx.range_pop()
        return {"W": w.squeeze(), "node_embedding": h, "edge_embedding": ea}


class PerfectEdgeClassification(nn.Module, HyperparametersMixin):
    def __init__(self, tpr=1.0, tnr=1.0, false_below_pt=0.0):
        """Truth-based edge "classifier" (reference edge_classifier.py:124-165): ``W`` is ``data.y``
        with true edges kept at rate ``tpr`` and false edges turned true at rate ``1 - tnr``, then
        everything with ``data.pt < false_below_pt`` set to false.  A few elementwise draws on the
        caller's device: no kernel of the path is involved."""
        super().__init__()
        self.save_hyperparameters()
        assert 0.0 <= tpr <= 1.0
        assert 0.0 <= tnr <= 1.0
        self.tpr, self.tnr, self.false_below_pt = tpr, tnr, false_below_pt

    def forward(self, data) -> dict[str, Tensor]:
        r = data.y.bool()
        if not math.isclose(self.tpr, 1.0):
            true_mask = r.detach().clone()
            r[true_mask] = torch.rand(int(true_mask.sum()), device=r.device) <= self.tpr
        if not math.isclose(self.tnr, 1.0):
            false_mask = (~r).detach().clone()
            r[false_mask] = ~(torch.rand(int(false_mask.sum()), device=r.device) <= self.tnr)
        if self.false_below_pt > 0.0:
            r[data.pt < self.false_below_pt] = False
        return {"W": r.float()}  # float like a trained classifier's output (and what BCE expects)
