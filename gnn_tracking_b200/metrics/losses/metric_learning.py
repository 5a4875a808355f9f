"""Hinge embedding loss for graph construction behind the reference interface (reference
metrics/losses/metric_learning.py:57-178): attraction over the true edges that start at a hit of
interest, repulsion over the radius-graph edges that start at a hit of interest and join different
particles.  The radius graph is never built: ``gtb_radius_pair_sum_f32`` walks the neighbourhoods
and sums the hinge terms in one pass (torch_cluster.radius_graph semantics, see include/gtb200.h)."""
from __future__ import annotations

import os

import torch
from torch import Tensor

from ... import ops
from ..._hparams import HyperparametersMixin
from ..._lib import check, lib
from ...utils.graph_masks import get_good_node_mask_tensors
from . import MultiLossFct, MultiLossFctReturn


def _grid_workspace(n: int, dev) -> tuple[Tensor, int]:
    ws_bytes = lib().gtb_radius_graph_grid_workspace_bytes(n)
    return torch.empty(ws_bytes, dtype=torch.uint8, device=dev), ws_bytes


def _use_grid() -> bool:
    """Neighbour search of the pair sums: the uniform cell list (``gtb_radius_pair_sum_grid_f32``) unless
    ``GTB_RADIUS_BRUTE=1`` asks for the all-pairs walk (``gtb_radius_pair_sum_f32``): same edges, same terms."""
    return os.environ.get("GTB_RADIUS_BRUTE") != "1"


class _RadiusPairSumFn(torch.autograd.Function):
    """``gtb_radius_pair_sum[_grid]_f32`` with the gradient of its first output (the sum of the pair terms)
    w.r.t. ``x`` and, in mode 1, ``beta`` (``gtb_radius_pair_sum_grad[_grid]_f32``); the third output (sum
    of ``beta`` over ``pid == 0``) is differentiated in place."""

    @staticmethod
    def forward(ctx, x, bt, pid, flag, b, cfg):
        q_min, r, p, eps, max_nb, mode = cfg
        dev = x.device
        n, d = x.shape
        out = torch.zeros(4, dtype=torch.float64, device=dev)
        args = (x.data_ptr(), d, n, None if b is None else b.data_ptr(), pid.data_ptr(), flag.data_ptr(),
                None if bt is None else bt.data_ptr(), q_min, r, p, eps, max_nb, mode, out.data_ptr())
        if _use_grid() and n > 0:
            ws, ws_bytes = _grid_workspace(n, dev)
            with ops.on_device(dev):
                check(lib().gtb_radius_pair_sum_grid_f32(*args, ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
            ops._count(6)  # cell list (bounding box, grid, keys, sort, layout) + the pair sums
        else:
            check(lib().gtb_radius_pair_sum_f32(*args, ops.stream_ptr(dev)))
            ops._count(1)
        ctx.save_for_backward(x, pid, flag, *([bt] if bt is not None else []), *([b] if b is not None else []))
        ctx.cfg, ctx.has = cfg, (bt is not None, b is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        q_min, r, p, eps, max_nb, mode = ctx.cfg
        has_bt, has_b = ctx.has
        saved = list(ctx.saved_tensors)
        x, pid, flag = saved[:3]
        bt = saved[3] if has_bt else None
        b = saved[3 + has_bt] if has_b else None
        n, d = x.shape
        dev = x.device
        coef = g[0:1].to(torch.float32).contiguous()
        gx = torch.zeros((n, d), dtype=torch.float32, device=dev)
        gq = torch.zeros(n, dtype=torch.float32, device=dev) if mode == 1 else None
        args = (x.data_ptr(), d, n, None if b is None else b.data_ptr(), pid.data_ptr(), flag.data_ptr(),
                None if bt is None else bt.data_ptr(), q_min, r, p, eps, max_nb, mode, coef.data_ptr(), gx.data_ptr(),
                None if gq is None else gq.data_ptr())
        if _use_grid() and n > 0:
            ws, ws_bytes = _grid_workspace(n, dev)
            with ops.on_device(dev):
                check(lib().gtb_radius_pair_sum_grad_grid_f32(*args, ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
            ops._count(6)
        else:
            check(lib().gtb_radius_pair_sum_grad_f32(*args, ops.stream_ptr(dev)))
            ops._count(1)
        gbeta = None
        if bt is not None:
            gbeta = (pid == 0).to(torch.float32) * g[2].to(torch.float32)
            if gq is not None:  # q = atanh(beta)^2 + q_min
                gbeta = gbeta + gq * 2.0 * torch.atanh(bt) / (1.0 - bt * bt)
        return gx, gbeta, None, None, None, None


def radius_pair_sum(*, x: Tensor, particle_id: Tensor, src_flag: Tensor, r: float, mode: int, batch: Tensor | None = None,
                    beta: Tensor | None = None, q_min: float = 0.0, p: float = 1.0, eps: float = 1e-9,
                    max_num_neighbors: int = 256) -> Tensor:
    """float64 [4]: {sum of terms, kept edges, sum of beta over pid == 0, hits with pid == 0}.
    Differentiable w.r.t. ``x`` and ``beta`` (entries 0 and 2)."""
    ops.require_cuda(x, particle_id, src_flag)
    x = x.to(torch.float32).contiguous()
    pid = particle_id.to(torch.int64).contiguous()
    flag = src_flag.to(torch.bool).contiguous().view(torch.uint8)
    b = None if batch is None else batch.to(torch.int64).contiguous()
    bt = None if beta is None else beta.reshape(-1).to(torch.float32).contiguous()
    cfg = (float(q_min), float(r), float(p), float(eps), int(max_num_neighbors), int(mode))
    return _RadiusPairSumFn.apply(x, bt, pid, flag, b, cfg)


class _EdgeDistPowSumFn(torch.autograd.Function):
    """``gtb_edge_dist_pow_sum_f32`` -> float64 {sum of dist^p over the flagged true edges, their number}."""

    @staticmethod
    def forward(ctx, x, tei, flag, p):
        dev = x.device
        att = torch.zeros(2, dtype=torch.float64, device=dev)
        check(lib().gtb_edge_dist_pow_sum_f32(x.data_ptr(), x.size(1), tei.data_ptr(), tei.size(1), flag.data_ptr(), float(p),
                                              att.data_ptr(), ops.stream_ptr(dev)))
        ops._count(1)
        ctx.save_for_backward(x, tei, flag)
        ctx.p = float(p)
        return att

    @staticmethod
    def backward(ctx, g):
        x, tei, flag = ctx.saved_tensors
        gx = torch.zeros_like(x)
        coef = g[0:1].to(torch.float32).contiguous()
        check(lib().gtb_edge_dist_pow_grad_f32(x.data_ptr(), x.size(1), tei.data_ptr(), tei.size(1), flag.data_ptr(), ctx.p,
                                               coef.data_ptr(), gx.data_ptr(), ops.stream_ptr(x.device)))
        ops._count(1)
        return gx, None, None, None


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
        ops.require_cuda(x, particle_id, true_edge_index)
        mask = get_good_node_mask_tensors(pt=pt, particle_id=particle_id, reconstructable=reconstructable, eta=eta,
                                          pt_thld=hp.pt_thld, max_eta=hp.max_eta)
        n_hits_oi = mask.sum()
        xf = x.to(torch.float32).contiguous()
        flag = mask.contiguous().view(torch.uint8)
        # attraction: true edges starting at a hit of interest (:111)
        tei = true_edge_index.to(torch.int64).contiguous()
        att = _EdgeDistPowSumFn.apply(xf, tei, flag, hp.p_attr)
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
