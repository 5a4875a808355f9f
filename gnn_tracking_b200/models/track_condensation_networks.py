"""Graph track-condensation network behind the reference interface (reference
models/track_condensation_networks.py:118-386): edge classifier -> threshold ->
edge sub-graph (-> orphan pruning) -> HC encoders -> ``hc_in`` ResIN -> beta / H heads.

The edge sub-graph is never materialised: the parent plan is stream-compacted
(``GraphPlan.filtered``) and the HC edge encoder gathers the kept rows of the parent
edge tensors directly."""
from __future__ import annotations

import importlib
import math

import torch
from torch import Tensor, nn

from .. import ops
from .._hparams import HyperparametersMixin
from ..ops import ACT_RELU, ACT_SIGMOID_AFFINE, Block
from ..plan import build_plan, get_plan, prune_orphans
from .edge_classifier import ECForGraphTCN, PerfectEdgeClassification
from .mlp import MLP, ResFCNN
from .resin import ResIN


def _obj_from_or_to_hparams(owner: HyperparametersMixin, key: str, obj):
    """The reference's hparams round-trip for sub-modules (utils/lightning.py:59-80):
    a ``{"class_path", "init_args"}`` dict is instantiated, a module is recorded."""
    if isinstance(obj, dict) and "class_path" in obj and "init_args" in obj:
        owner.save_hyperparameters({key: obj})
        mod, _, name = obj["class_path"].rpartition(".")
        return getattr(importlib.import_module(mod), name)(**obj["init_args"])
    if obj is None or isinstance(obj, (int, float, str, bool, list, tuple, dict)):
        owner.save_hyperparameters({key: obj})
        return obj
    cls = type(obj)
    owner.save_hyperparameters({key: {"class_path": f"{cls.__module__}.{cls.__qualname__}",
                                      "init_args": dict(getattr(obj, "hparams", {}))}})
    return obj


class ModularGraphTCN(nn.Module, HyperparametersMixin):
    def __init__(self, *, ec: nn.Module | None = None, hc_in: nn.Module, node_indim: int, edge_indim: int,
                 h_dim: int = 5, e_dim: int = 4, h_outdim: int = 2, hidden_dim: int = 40,
                 feed_edge_weights: bool = False, ec_threshold: float = 0.5, mask_orphan_nodes: bool = False,
                 use_ec_embeddings_for_hc: bool = False, alpha_latent: float = 0.0, n_embedding_coords: int = 0,
                 heterogeneous_node_encoder: bool = False):
        super().__init__()
        self.save_hyperparameters(ignore=["ec", "hc_in"])
        if heterogeneous_node_encoder:
            raise NotImplementedError("heterogeneous_node_encoder (pixel/strip split) is outside the B200 hot path")
        self.relu = nn.ReLU()
        self.ec = _obj_from_or_to_hparams(self, "ec", ec)
        self.hc_in = _obj_from_or_to_hparams(self, "hc_in", hc_in)
        node_enc_indim, edge_enc_indim = node_indim, edge_indim
        if use_ec_embeddings_for_hc:
            ec_node_dim, ec_edge_dim = self.ec.latent_dim
            node_enc_indim += int(ec_node_dim)
            edge_enc_indim += int(ec_edge_dim)
        edge_enc_indim += int(feed_edge_weights)
        self.hc_edge_encoder = MLP(edge_enc_indim, e_dim, hidden_dim=hidden_dim, L=2, bias=False)
        self.hc_node_encoder = ResFCNN(in_dim=node_enc_indim, out_dim=h_dim, hidden_dim=hidden_dim, depth=1,
                                       bias=False, alpha=0)
        self.p_beta = MLP(h_dim, 1, hidden_dim, L=3)
        self.p_cluster = MLP(h_dim, h_outdim, hidden_dim, L=3)
        self._latent_normalization = nn.Parameter(torch.Tensor([1.0]), requires_grad=True)

    def forward(self, data) -> dict[str, Tensor | None]:
        hp = self.hparams
        x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        dev = ops.require_cuda(x, edge_index, edge_attr)
        n = x.size(0)
        plan = get_plan(edge_index, n)
        w_unmasked = edge_mask = hit_mask = None
        kept = None          # int32 ids of the surviving edges (None: all edges, in place)
        node_ids = None      # int32 ids of the surviving nodes (None: all nodes)
        ec_node_emb = ec_edge_emb = edge_weights = None
        if self.ec is not None:
            ec_out = self.ec(data)
            # the reference attaches the EC output to the caller's data object (:245-249)
            data.edge_weights = ec_out["W"].reshape((-1, 1))
            data.ec_node_embedding = ec_out.get("node_embedding", None)
            data.ec_edge_embedding = ec_out.get("edge_embedding", None)
            edge_weights, ec_node_emb, ec_edge_emb = data.edge_weights, data.ec_node_embedding, data.ec_edge_embedding
            w_unmasked = data.edge_weights.squeeze()
            edge_mask = (data.edge_weights > hp.ec_threshold).squeeze()
            plan, _, kept = plan.filtered(edge_mask.reshape(-1))
            if hp.mask_orphan_nodes:
                # unique endpoints of the surviving edges, relabelled in increasing order (:254-259)
                # (a compaction of the filtered plan: the relabelling is monotone, nothing is sorted again)
                plan, node_ids, new_id = prune_orphans(plan)
                hit_mask = new_id >= 0
                n = plan.n_nodes
            else:
                hit_mask = torch.ones(n, dtype=torch.bool, device=dev)
        elif hp.feed_edge_weights:
            data.edge_weights = data.ec_score.reshape((-1, 1))
            edge_weights = data.edge_weights
        e = plan.n_edges

        # encoder inputs: column blocks gathered straight from the parent tensors (:268-280)
        xs = [Block(x, node_ids)]
        es = [Block(edge_attr, kept)]
        if hp.use_ec_embeddings_for_hc:
            assert ec_edge_emb is not None and ec_node_emb is not None
            xs.append(Block(ec_node_emb, node_ids))
            es.append(Block(ec_edge_emb, kept))
        if hp.feed_edge_weights:
            es.append(Block(edge_weights, kept))
        h = self.hc_node_encoder.forward_blocks(xs, n, final_act=ACT_RELU)
        ea = self.hc_edge_encoder.forward_blocks(es, e, final_act=ACT_RELU)

        h, _, _ = self.hc_in.forward_planned(h, plan, ea)
        beta = self.p_beta.forward_blocks([Block(h)], n, final_act=ACT_SIGMOID_AFFINE, act_eps=1e-6)
        epi = {}
        if alpha_residue := hp.alpha_latent:
            nec: int = hp.n_embedding_coords
            assert 0 < nec <= hp.h_outdim
            xr = x if node_ids is None else x[node_ids.long()]
            epi = dict(res=nn.functional.pad(xr[:, :nec], (0, hp.h_outdim - nec)).contiguous(),
                       res_a=math.sqrt(alpha_residue), res_b=math.sqrt(1 - alpha_residue))
        # H = (sqrt(a) residual + sqrt(1-a) p_cluster(h)) * _latent_normalization  (:290-298)
        if torch.is_grad_enabled() and (h.requires_grad or self._latent_normalization.requires_grad):
            # the scale is a trainable scalar: its gradient comes from one elementwise product
            hh = self.p_cluster.forward_blocks([Block(h)], n, **epi) * self._latent_normalization
        else:
            hh = self.p_cluster.forward_blocks([Block(h)], n, out_scale=self._latent_normalization.detach(), **epi)
        return {"W": w_unmasked, "H": hh, "B": beta.squeeze(), "ec_hit_mask": hit_mask, "ec_edge_mask": edge_mask}


class GraphTCN(nn.Module, HyperparametersMixin):
    def __init__(self, node_indim: int, edge_indim: int, *, h_dim=5, e_dim=4, h_outdim=2, hidden_dim=40,
                 L_ec=3, L_hc=3, alpha_ec: float = 0.5, alpha_hc: float = 0.5, **kwargs):
        """``ModularGraphTCN`` with an ``ECForGraphTCN`` edge classifier and a ``ResIN``
        track condenser (reference :311-386); parameters under ``_gtcn.*``."""
        super().__init__()
        self.save_hyperparameters()
        ec = ECForGraphTCN(node_indim=node_indim, edge_indim=edge_indim, hidden_dim=hidden_dim,
                           interaction_node_dim=h_dim, interaction_edge_dim=e_dim, L_ec=L_ec, alpha=alpha_ec)
        hc_in = ResIN(node_dim=h_dim, edge_dim=e_dim, object_hidden_dim=hidden_dim,
                      relational_hidden_dim=hidden_dim, alpha=alpha_hc, n_layers=L_hc)
        self._gtcn = ModularGraphTCN(ec=ec, hc_in=hc_in, node_indim=node_indim, edge_indim=edge_indim,
                                     h_dim=h_dim, e_dim=e_dim, h_outdim=h_outdim, hidden_dim=hidden_dim, **kwargs)

    def forward(self, data) -> dict[str, Tensor | None]:
        return self._gtcn.forward(data=data)


class PreTrainedECGraphTCN(nn.Module, HyperparametersMixin):
    def __init__(self, ec, *, node_indim: int, edge_indim: int, h_dim=5, e_dim=4, h_outdim=2, hidden_dim=40,
                 L_hc=3, alpha_hc: float = 0.5, **kwargs):
        """``ModularGraphTCN`` around a given (pre-trained) edge classifier ``ec`` — a module or a
        ``{"class_path", "init_args"}`` dict — with a ``ResIN`` track condenser (reference :459-518,
        the model of tests/test_configs/tc.yml); parameters under ``_gtcn.*``."""
        super().__init__()
        self.save_hyperparameters(ignore=["ec"])
        ec = _obj_from_or_to_hparams(self, "ec", ec)
        hc_in = ResIN(node_dim=h_dim, edge_dim=e_dim, object_hidden_dim=hidden_dim,
                      relational_hidden_dim=hidden_dim, alpha=alpha_hc, n_layers=L_hc)
        self._gtcn = ModularGraphTCN(ec=ec, hc_in=hc_in, node_indim=node_indim, edge_indim=edge_indim,
                                     h_dim=h_dim, e_dim=e_dim, h_outdim=h_outdim, hidden_dim=hidden_dim, **kwargs)

    def forward(self, data) -> dict[str, Tensor | None]:
        return self._gtcn.forward(data=data)


class PerfectECGraphTCN(nn.Module, HyperparametersMixin):
    def __init__(self, *, node_indim: int, edge_indim: int, h_dim=5, e_dim=4, h_outdim=2, hidden_dim=40, L_hc=3,
                 alpha_hc: float = 0.5, ec_tpr=1.0, ec_tnr=1.0, **kwargs):
        """``ModularGraphTCN`` behind a truth-based edge classifier with the given true-positive /
        true-negative rates (reference :389-456)."""
        super().__init__()
        self.save_hyperparameters()
        ec = PerfectEdgeClassification(tpr=ec_tpr, tnr=ec_tnr)
        hc_in = ResIN(node_dim=h_dim, edge_dim=e_dim, object_hidden_dim=hidden_dim,
                      relational_hidden_dim=hidden_dim, alpha=alpha_hc, n_layers=L_hc)
        self._gtcn = ModularGraphTCN(ec=ec, hc_in=hc_in, node_indim=node_indim, edge_indim=edge_indim,
                                     h_dim=h_dim, e_dim=e_dim, h_outdim=h_outdim, hidden_dim=hidden_dim, **kwargs)

    def forward(self, data) -> dict[str, Tensor | None]:
        return self._gtcn.forward(data=data)
