"""Hinge embedding loss for graph construction behind the reference interface (reference
metrics/losses/metric_learning.py:57-178): attraction over the true edges that start at a hit of
interest, repulsion over the radius-graph edges that start at a hit of interest and join different
particles.  The radius graph is never built: ``gtb_radius_pair_sum_f32`` walks the neighbourhoods
and sums the hinge terms in one pass (torch_cluster.radius_graph semantics, see include/gtb200.h)."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import ops
from ..._hparams import HyperparametersMixin
from ..._lib import check, lib
from ...utils.graph_masks import get_good_node_mask_tensors
from . import MultiLossFct, MultiLossFctReturn


def radius_pair_sum(*, x: Tensor, particle_id: Tensor, src_flag: Tensor, r: float, mode: int, batch: Tensor | None = None,
                    beta: Tensor | None = None, q_min: float = 0.0, p: float = 1.0, eps: float = 1e-9,
                    max_num_neighbors: int = 256) -> Tensor:
    """float64 [4]: {sum of terms, kept edges, sum of beta over pid == 0, hits with pid == 0}."""
    dev = ops.require_cuda(x, particle_id, src_flag)
    if torch.is_grad_enabled() and (x.requires_grad or (beta is not None and beta.requires_grad)):
        raise NotImplementedError("the radius-graph losses are forward-only in this build: call them under torch.no_grad()")
    x = x.to(torch.float32).contiguous()
    n, d = x.shape
    pid = particle_id.to(torch.int64).contiguous()
    flag = src_flag.to(torch.bool).contiguous().view(torch.uint8)
    b = None if batch is None else batch.to(torch.int64).contiguous()
    bt = None if beta is None else beta.reshape(-1).to(torch.float32).contiguous()
    out = torch.zeros(4, dtype=torch.float64, device=dev)
    check(lib().gtb_radius_pair_sum_f32(x.data_ptr(), d, n, None if b is None else b.data_ptr(), pid.data_ptr(), flag.data_ptr(),
                                        None if bt is None else bt.data_ptr(), float(q_min), float(r), float(p), float(eps),
                                        int(max_num_neighbors), int(mode), out.data_ptr(), ops.stream_ptr(dev)))
    ops._count(1)
    return out


class GraphConstructionHingeEmbeddingLoss(MultiLossFct, HyperparametersMixin):
    def __init__(self, *, lw_repulsive: float = 1.0, r_emb: float = 1.0, max_num_neighbors: int = 256,
                 pt_thld: float = 0.9, max_eta: float = 4.0, p_attr: float = 1.0, p_rep: float = 1.0,
                 rep_normalization: str = "n_hits_oi", rep_oi_only: bool = True):
        """Same arguments as the reference (metric_learning.py:59-91)."""
        super().__init__()
        self.save_hyperparameters()
        if rep_normalization not in ("n_rep_edges", "n_hits_oi", "n_att_edges"):
            raise ValueError(f"Normalization {rep_normalization} not recognized.")

    def forward(self, *, x: Tensor, particle_id: Tensor, batch: Tensor, true_edge_index: Tensor, pt: Tensor,
                eta: Tensor, reconstructable: Tensor, **kwargs) -> MultiLossFctReturn:
        hp = self.hparams
        if true_edge_index is None:
            raise ValueError("True_edge_index must be given and not be None. Are you trying to use this loss for OC "
                             "training? In this case, double check that you are properly passing on the true edges.")
        dev = ops.require_cuda(x, particle_id, true_edge_index)
        mask = get_good_node_mask_tensors(pt=pt, particle_id=particle_id, reconstructable=reconstructable, eta=eta,
                                          pt_thld=hp.pt_thld, max_eta=hp.max_eta)
        n_hits_oi = mask.sum()
        xf = x.to(torch.float32).contiguous()
        flag = mask.contiguous().view(torch.uint8)
        # attraction: true edges starting at a hit of interest (:111)
        tei = true_edge_index.to(torch.int64).contiguous()
        att = torch.zeros(2, dtype=torch.float64, device=dev)
        check(lib().gtb_edge_dist_pow_sum_f32(xf.data_ptr(), xf.size(1), tei.data_ptr(), tei.size(1), flag.data_ptr(),
                                              float(hp.p_attr), att.data_ptr(), ops.stream_ptr(dev)))
        ops._count(1)
        # repulsion: radius-graph edges starting at a hit of interest (or any hit), different particles (:97-110)
        src_flag = mask if hp.rep_oi_only else torch.ones_like(mask)
        rep = radius_pair_sum(x=xf, particle_id=particle_id, src_flag=src_flag, r=hp.r_emb, mode=0, batch=batch,
                              p=hp.p_rep, max_num_neighbors=hp.max_num_neighbors)
        eps = 1e-9
        norm_rep = {"n_rep_edges": rep[1], "n_hits_oi": n_hits_oi.to(torch.float64), "n_att_edges": att[1]}[hp.rep_normalization] + eps
        losses = {"attractive": (att[0] / (att[1] + eps)).float(), "repulsive": (rep[0] / norm_rep).float()}
        weights = {"attractive": 1.0, "repulsive": hp.lw_repulsive}
        extra = {"n_hits_oi": n_hits_oi, "n_edges_att": att[1].to(torch.int64), "n_edges_rep": rep[1].to(torch.int64)}
        return MultiLossFctReturn(loss_dct=losses, weight_dct=weights, extra_metrics=extra)
