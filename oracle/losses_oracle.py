"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the losses on the hot path
(reference ``metrics/losses/{ec,oc,metric_learning}.py``, ``utils/graph_masks.py``).

Parity status: PINNED against the reference's own known-answer values in
``/root/reference/tests/test_losses.py:112-123`` (condensation losses, tiger == RG)
and ``:194-203`` (hinge), reproduced in ``tests/test_oracle_losses.py`` from the
seeded generator restated in ``tests/golden/loss_testdata.py``.

``file:line`` citations are relative to ``/root/reference/src/gnn_tracking``.
Only tests / smoke / the bench CPU legs may import this module.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor


def good_node_mask(*, pt, particle_id, reconstructable, eta, pt_thld=0.9, max_eta=4.0) -> Tensor:
    """``get_good_node_mask_tensors`` utils/graph_masks.py:19-28."""
    return (pt > pt_thld) & (particle_id > 0) & (reconstructable > 0) & (eta.abs() < max_eta)


# ------------------------------------------------------------------------ EC losses
def falsify_low_pt_edges(*, y, edge_index=None, pt=None, pt_thld=0.0):
    """metrics/losses/ec.py:71-92 -- only the pt of edge_index[0] is looked at."""
    if math.isclose(pt_thld, 0.0):
        return y
    return y.bool() & (pt[edge_index[0, :]] > pt_thld)


def bce_mean(w: Tensor, y: Tensor) -> Tensor:
    """``EdgeWeightBCELoss._forward`` metrics/losses/ec.py:116-121 =
    ``F.binary_cross_entropy(w, y, 'mean')``; torch clamps each log at -100."""
    lw = torch.log(w).clamp_min(-100.0)
    l1w = torch.log(1 - w).clamp_min(-100.0)
    return (-(y * lw + (1 - y) * l1w)).mean()


def focal_mean(w: Tensor, y: Tensor, *, alpha=0.25, gamma=2.0, pos_weight: Tensor | None = None) -> Tensor:
    """``_binary_focal_loss`` metrics/losses/ec.py:12-29."""
    if pos_weight is None:
        pos_weight = torch.tensor([1.0], dtype=w.dtype)
    pn = 1 - w
    pos = -alpha * pos_weight * pn.pow(gamma) * y * w.log()
    neg = -(1.0 - alpha) * w.pow(gamma) * (1.0 - y) * pn.log()
    return (pos + neg).mean()


def edge_weight_bce(*, w, y, edge_index=None, pt=None, pt_thld=0.0):
    """``FalsifyLowPtEdgeWeightLoss.forward`` ec.py:103-109 + BCE."""
    y = falsify_low_pt_edges(y=y, edge_index=edge_index, pt=pt, pt_thld=pt_thld)
    return bce_mean(w, y.to(w.dtype))


def edge_weight_focal(*, w, y, edge_index=None, pt=None, pt_thld=0.0, alpha=0.25, gamma=2.0, pos_weight=None):
    """``EdgeWeightFocalLoss`` ec.py:124-150."""
    y = falsify_low_pt_edges(y=y, edge_index=edge_index, pt=pt, pt_thld=pt_thld)
    return focal_mean(w, y.to(w.dtype), alpha=alpha, gamma=gamma, pos_weight=pos_weight)


def haughty_focal(*, w, y, edge_index, pt, pt_thld=0.0, alpha=0.25, gamma=2.0):
    """``HaughtyFocalLoss.forward`` ec.py:167-178: falsified labels are the
    ``pos_weight``; the raw ``y.long()`` is the target."""
    pw = falsify_low_pt_edges(y=y, edge_index=edge_index, pt=pt, pt_thld=pt_thld)
    return focal_mean(w, y.long(), alpha=alpha, gamma=gamma, pos_weight=pw)


# --------------------------------------------------------------- condensation tiger
def condensation_tiger(*, beta, x, object_id, object_mask, q_min=0.01, noise_threshold=0):
    """``condensation_loss_tiger`` metrics/losses/oc.py:251-347 with sampling off
    (``max_n_rep=0``).  Dense N x K formulation, exactly the reference's order of
    operations (cdist, masks, sums)."""
    eps = 1e-9
    uniq = torch.unique(object_id[object_mask])
    att = object_id.view(-1, 1) == uniq.view(1, -1)
    q = torch.arctanh(beta) ** 2 + q_min
    alphas = torch.argmax(q.view(-1, 1) * att, dim=0)
    qk = q[alphas].view(1, -1)
    qw = q.view(-1, 1) * qk
    xk = x[alphas]
    dist = torch.cdist(x, xk)
    n_hits = len(object_mask)
    n_hits_oi = object_mask.sum()
    k = len(alphas)
    norm_rep = eps + (k - 1) * n_hits
    norm_att = eps + n_hits_oi - k
    v_att = (qw[att] * dist[att].square()).sum() / norm_att
    rep = (~att) & (dist < 1)
    n_rep = rep.sum()
    v_rep = (qw[rep] * (1 - dist[rep])).sum() / norm_rep
    coward = torch.mean(1 - beta[alphas])
    noise = torch.mean(beta[~(object_id > noise_threshold)])
    return ({"attractive": v_att, "repulsive": v_rep, "coward": coward, "noise": noise},
            {"n_rep": n_rep, "alphas": alphas, "unique_ids": uniq})


def condensation_tiger_loss(*, beta, x, particle_id, reconstructable, pt, eta, ec_hit_mask=None,
                            q_min=0.01, pt_thld=0.9, max_eta=4.0):
    """``CondensationLossTiger.forward`` oc.py:382-436 (sample_pids = 1)."""
    if ec_hit_mask is not None:
        particle_id, reconstructable, pt, eta = (t[ec_hit_mask] for t in (particle_id, reconstructable, pt, eta))
    mask = good_node_mask(pt=pt, particle_id=particle_id, reconstructable=reconstructable, eta=eta,
                          pt_thld=pt_thld, max_eta=max_eta)
    return condensation_tiger(beta=beta, x=x, object_id=particle_id, object_mask=mask, q_min=q_min)


# ------------------------------------------------------------- radius graph + RG OC
def radius_graph(x: Tensor, r: float, batch: Tensor | None = None, max_num_neighbors: int | None = None) -> Tensor:
    """``torch_cluster.radius_graph(x, r, batch, loop=False, max_num_neighbors)``
    restated (un-vendored dependency, reference ``environments/default.yml:15``;
    call sites oc.py:115-117, metric_learning.py:97-103): strict ``dist < r``, no
    self loops, same ``batch`` only; row 0 = neighbour, row 1 = centre.  When a
    centre has more than ``max_num_neighbors`` neighbours torch_cluster keeps an
    implementation-defined subset; this oracle keeps the lowest indices."""
    d2 = ((x.unsqueeze(1) - x.unsqueeze(0)) ** 2).sum(-1)
    adj = d2 < r * r
    if batch is not None:
        adj &= batch.view(-1, 1) == batch.view(1, -1)
    adj.fill_diagonal_(False)
    centre, neigh = adj.nonzero(as_tuple=True)
    if max_num_neighbors is not None and centre.numel():
        start = torch.ones_like(centre, dtype=torch.bool)
        start[1:] = centre[1:] != centre[:-1]
        first = torch.where(start)[0]
        rank = torch.arange(centre.numel()) - first[torch.cumsum(start.long(), 0) - 1]
        keep = rank < max_num_neighbors
        centre, neigh = centre[keep], neigh[keep]
    return torch.stack([neigh, centre])


def first_occurrences(x: Tensor) -> Tensor:
    """``_first_occurrences`` oc.py:16-23: index of the first occurrence of each
    unique value (sorted by value)."""
    uniq, inv = torch.unique(x, sorted=True, return_inverse=True)
    out = torch.full((uniq.numel(),), x.numel(), dtype=torch.long)
    out.scatter_reduce_(0, inv, torch.arange(x.numel()), reduce="amin")
    return out


def condensation_rg(*, beta, x, particle_id, mask, q_min=0.01, radius_threshold=1.0, max_num_neighbors=256):
    """``_radius_graph_condensation_loss`` oc.py:87-161 (+ helpers :32-84)."""
    order = torch.argsort(beta[mask], descending=True)
    pids_sorted = particle_id[mask][order]
    alphas_masked = order[first_occurrences(pids_sorted)]
    alphas = torch.nonzero(mask).squeeze()[alphas_masked]
    is_cp = torch.zeros_like(particle_id, dtype=torch.bool)
    is_cp[alphas] = True
    q = torch.arctanh(beta) ** 2 + q_min
    re = radius_graph(x, radius_threshold, max_num_neighbors=max_num_neighbors)
    eps = 1e-9
    sel = is_cp[re[0]] & (particle_id[re[0]] != particle_id[re[1]])
    rep = re[:, sel]
    d = radius_threshold - torch.sqrt(eps + ((x[rep[0]] - x[rep[1]]) ** 2).sum(-1))
    vr = (d * q[rep[0]] * q[rep[1]]).sum()
    non_cp = torch.nonzero(~is_cp & mask).squeeze()
    corr = alphas[torch.searchsorted(particle_id[alphas], particle_id[non_cp])]
    va = (((x[non_cp] - x[corr]) ** 2).sum(-1) * q[non_cp] * q[corr]).sum()
    n_hits = len(mask)
    k = len(alphas)
    norm_rep = eps + (k - 1) * n_hits
    norm_att = eps + mask.sum() - k
    return {
        "attractive": va / norm_att,
        "repulsive": vr / norm_rep,
        "coward": torch.mean(1 - beta[alphas]),
        "noise": torch.mean(beta[particle_id == 0]),
    }


# ----------------------------------------------------------------------- hinge loss
def hinge_loss(*, x, particle_id, batch, true_edge_index, pt, eta, reconstructable, r_emb=1.0,
               max_num_neighbors=256, pt_thld=0.9, max_eta=4.0, p_attr=1.0, p_rep=1.0,
               rep_normalization="n_hits_oi", rep_oi_only=True):
    """``GraphConstructionHingeEmbeddingLoss.forward`` metric_learning.py:114-178
    with ``_get_edges`` :93-112 and ``_hinge_loss_components`` :14-54."""
    mask = good_node_mask(pt=pt, particle_id=particle_id, reconstructable=reconstructable, eta=eta,
                          pt_thld=pt_thld, max_eta=max_eta)
    n_hits_oi = int(mask.sum())
    near = radius_graph(x, r_emb, batch=batch, max_num_neighbors=max_num_neighbors)
    rep = near[:, mask[near[0]]] if rep_oi_only else near
    rep = rep[:, particle_id[rep[0]] != particle_id[rep[1]]]
    att = true_edge_index[:, mask[true_edge_index[0]]]
    eps = 1e-9
    d_att = (x[att[0]] - x[att[1]]).norm(dim=-1)
    v_att = d_att.pow(p_attr).sum() / (att.shape[1] + eps)
    d_rep = (x[rep[0]] - x[rep[1]]).norm(dim=-1)
    if rep_normalization == "n_rep_edges":
        nr = rep.shape[1] + eps
    elif rep_normalization == "n_hits_oi":
        nr = n_hits_oi + eps
    elif rep_normalization == "n_att_edges":
        nr = att.shape[1] + eps
    else:
        raise ValueError(rep_normalization)
    v_rep = torch.relu(r_emb - d_rep.pow(p_rep)).sum() / nr
    return ({"attractive": v_att, "repulsive": v_rep},
            {"n_hits_oi": n_hits_oi, "n_edges_att": att.shape[1], "n_edges_rep": rep.shape[1]})
