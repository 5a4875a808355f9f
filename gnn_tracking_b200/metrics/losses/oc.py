"""Object-condensation losses behind the reference interface: "tiger" (reference
metrics/losses/oc.py:251-436) without the N x K planes (``gtb_oc_*``) and the radius-graph variant
(oc.py:87-248) without the radius graph (``gtb_radius_pair_sum_f32``)."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import ops
from ..._hparams import HyperparametersMixin
from ..._lib import check, lib
from ...utils.graph_masks import get_good_node_mask_tensors
from . import MultiLossFct, MultiLossFctReturn


class _TigerFn(torch.autograd.Function):
    """The four tiger loss terms with their analytic gradients w.r.t. ``beta`` and ``x``
    (``gtb_oc_potentials_grad``)."""

    @staticmethod
    def forward(ctx, beta, x, oid, mask, q_min, noise_threshold):
        dev = beta.device
        n, d = x.shape
        st = ops.stream_ptr(dev)
        uniq = torch.empty(n, dtype=torch.int64, device=dev)
        slot = torch.empty(n, dtype=torch.int32, device=dev)
        n_uniq = torch.zeros(1, dtype=torch.int32, device=dev)
        ws_bytes = lib().gtb_oc_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().gtb_oc_prepare(oid.data_ptr(), mask.data_ptr(), n, uniq.data_ptr(), slot.data_ptr(),
                                   n_uniq.data_ptr(), ws.data_ptr(), ws_bytes, st))
        k = int(n_uniq.item())  # the reference's torch.unique syncs at the same place
        assert k > 0, "No hits left after masking"
        scratch = torch.empty(k, dtype=torch.int64, device=dev)
        alphas = torch.empty(k, dtype=torch.int32, device=dev)
        check(lib().gtb_oc_alphas(beta.data_ptr(), slot.data_ptr(), n, float(q_min), k, scratch.data_ptr(),
                                  alphas.data_ptr(), st))
        out = torch.zeros(8, dtype=torch.float64, device=dev)
        check(lib().gtb_oc_potentials(beta.data_ptr(), x.data_ptr(), d, oid.data_ptr(), mask.data_ptr(),
                                      slot.data_ptr(), n, alphas.data_ptr(), k, float(q_min), int(noise_threshold),
                                      out.data_ptr(), st))
        ops._count(9)
        eps = 1e-9
        v_att, v_rep, coward, noise, n_noise, n_oi, n_rep, _ = out.unbind(0)
        norms = torch.stack([eps + n_oi - k, (eps + (k - 1) * n) * torch.ones_like(n_oi), k * torch.ones_like(n_oi), n_noise])
        ctx.save_for_backward(beta, x, oid, slot, alphas, norms)
        ctx.cfg = (float(q_min), int(noise_threshold), k)
        # noise: NaN without noise hits, as in the reference (oc.py:335-336)
        return ((v_att / norms[0]).float(), (v_rep / norms[1]).float(), (coward / k).float(), (noise / n_noise).float(),
                n_rep.to(torch.int64), alphas, uniq[:k])

    @staticmethod
    def backward(ctx, g_att, g_rep, g_cow, g_noise, *unused):
        beta, x, oid, slot, alphas, norms = ctx.saved_tensors
        q_min, noise_threshold, k = ctx.cfg
        n, d = x.shape
        dev = x.device
        g = torch.stack([t.to(torch.float64).reshape(()) for t in (g_att, g_rep, g_cow, g_noise)])
        coef = (g / norms).to(torch.float32).contiguous()
        gq = torch.empty(n, dtype=torch.float32, device=dev)
        gbeta = torch.empty(n, dtype=torch.float32, device=dev)
        gx = torch.empty((n, d), dtype=torch.float32, device=dev)
        check(lib().gtb_oc_potentials_grad(beta.data_ptr(), x.data_ptr(), d, oid.data_ptr(), slot.data_ptr(), n,
                                           alphas.data_ptr(), k, q_min, noise_threshold, coef.data_ptr(), gq.data_ptr(),
                                           gbeta.data_ptr(), gx.data_ptr(), ops.stream_ptr(dev)))
        ops._count(3)
        return gbeta, gx, None, None, None, None


def condensation_loss_tiger(*, beta: Tensor, x: Tensor, object_id: Tensor, object_mask: Tensor, q_min: float,
                            noise_threshold: int = 0, max_n_rep: int = 0):
    """Same contract as the reference function (oc.py:251-347): returns
    ``({"attractive", "repulsive", "coward", "noise"}, {"n_rep"})``.  Differentiable w.r.t. ``beta``
    and ``x``."""
    # max_n_rep (oc.py:320-328): the reference keeps each repulsive pair with probability max_n_rep / n_rep and
    # divides the normalisation by the same factor -- an unbiased estimate of the full sum that exists to bound
    # the memory of its N x K planes.  The tiled kernel has no such planes: it returns the full sum, i.e. the
    # expectation of the reference's estimate, without the sampling noise.
    del max_n_rep
    ops.require_cuda(beta, x, object_id, object_mask)
    shape = beta.shape
    beta = beta.reshape(-1).to(torch.float32).contiguous()
    x = x.to(torch.float32).contiguous()
    oid = object_id.to(torch.int64).contiguous()
    mask = object_mask.to(torch.bool).contiguous().view(torch.uint8)
    del shape
    att, rep, cow, noise, n_rep, alphas, uniq = _TigerFn.apply(beta, x, oid, mask, q_min, noise_threshold)
    losses = {"attractive": att, "repulsive": rep, "coward": cow, "noise": noise}
    return losses, {"n_rep": n_rep, "alphas": alphas, "unique_ids": uniq}


def condensation_loss_rg(*, beta: Tensor, x: Tensor, particle_id: Tensor, mask: Tensor, q_min: float,
                         radius_threshold: float = 1.0, max_num_neighbors: int = 256):
    """``_radius_graph_condensation_loss`` (oc.py:87-161).  The reference picks its condensation points
    among the MASKED hits only (argsort of ``beta[mask]`` + first occurrence per particle, oc.py:33-43),
    attracts the masked non-CP hits only (oc.py:75-84) and normalises with ``mask.sum()`` -- unlike the
    tiger loss, whose objects own every hit carrying their id.  The tiger kernels give exactly that when
    the hits outside the mask are taken out of their particle (object id -1: never an object of interest,
    masked hits have ids > 0).  The repulsion runs over the radius-graph edges that start at a
    condensation point (``gtb_radius_pair_sum_f32`` mode 1: ``sqrt(1e-9 + d^2)``, neighbour cap, the
    ORIGINAL particle ids decide which pairs repel), the noise term over ``particle_id == 0`` exactly."""
    from .metric_learning import radius_pair_sum
    mask = mask.to(torch.bool)
    masked_id = torch.where(mask, particle_id, torch.full_like(particle_id, -1))
    tiger, extra = condensation_loss_tiger(beta=beta, x=x, object_id=masked_id, object_mask=mask, q_min=q_min)
    n = x.size(0)
    k = extra["alphas"].numel()
    is_cp = torch.zeros(n, dtype=torch.bool, device=x.device)
    is_cp[extra["alphas"].long()] = True
    rep = radius_pair_sum(x=x, particle_id=particle_id, src_flag=is_cp, r=radius_threshold, mode=1, beta=beta, q_min=q_min,
                          eps=1e-9, max_num_neighbors=max_num_neighbors)
    eps = 1e-9
    losses = {
        "attractive": tiger["attractive"],
        "repulsive": (rep[0] / (eps + (k - 1) * n)).float(),
        "coward": tiger["coward"],
        "noise": (rep[2] / rep[3]).float(),  # NaN without noise hits, as torch.mean of an empty selection
    }
    return losses, {}


class CondensationLossTiger(MultiLossFct, HyperparametersMixin):
    def __init__(self, *, lw_repulsive: float = 1.0, lw_noise: float = 0.0, lw_coward: float = 0.0,
                 q_min: float = 0.01, pt_thld: float = 0.9, max_eta: float = 4.0, max_n_rep: int = 0,
                 sample_pids: float = 1.0):
        super().__init__()
        self.save_hyperparameters()

    def forward(self, *, beta: Tensor, x: Tensor, particle_id: Tensor, reconstructable: Tensor, pt: Tensor,
                ec_hit_mask: Tensor | None = None, eta: Tensor, **kwargs) -> MultiLossFctReturn:
        hp = self.hparams
        if ec_hit_mask is not None:  # model outputs are already pruned, the truth is not (oc.py:394-401)
            particle_id, reconstructable, pt, eta = (t[ec_hit_mask] for t in (particle_id, reconstructable, pt, eta))
        mask = get_good_node_mask_tensors(pt=pt, particle_id=particle_id, reconstructable=reconstructable,
                                          eta=eta, pt_thld=hp.pt_thld, max_eta=hp.max_eta)
        if hp.sample_pids < 1:  # the reference's own draw (oc.py:410-414): same call, same torch RNG stream
            mask = mask & (torch.rand_like(beta, dtype=torch.float16) < hp.sample_pids)
        assert mask.sum() > 0, "No hits left after masking"
        losses, extra = condensation_loss_tiger(beta=beta, x=x, object_id=particle_id, object_mask=mask,
                                                q_min=hp.q_min, noise_threshold=0, max_n_rep=hp.max_n_rep)
        weights = {"attractive": 1.0, "repulsive": hp.lw_repulsive, "noise": hp.lw_noise, "coward": hp.lw_coward}
        return MultiLossFctReturn(loss_dct=losses, weight_dct=weights, extra_metrics={"n_rep": extra["n_rep"]})


class CondensationLossRG(MultiLossFct, HyperparametersMixin):
    def __init__(self, *, lw_repulsive: float = 1.0, lw_noise: float = 0.0, lw_coward: float = 0.0,
                 q_min: float = 0.01, pt_thld: float = 0.9, max_eta: float = 4.0, max_num_neighbors: int = 256,
                 sample_pids: float = 1.0):
        """Radius-graph condensation loss with the reference's arguments (oc.py:164-194)."""
        super().__init__()
        self.save_hyperparameters()

    def forward(self, *, beta: Tensor, x: Tensor, particle_id: Tensor, reconstructable: Tensor, pt: Tensor,
                ec_hit_mask: Tensor | None = None, eta: Tensor, **kwargs) -> MultiLossFctReturn:
        hp = self.hparams
        if ec_hit_mask is not None:
            # the reference masks particle_id / reconstructable / pt but forgets eta (oc.py:207-213), which
            # makes its mask computation fail on pruned graphs; eta is masked here as the tiger loss does
            particle_id, reconstructable, pt, eta = (t[ec_hit_mask] for t in (particle_id, reconstructable, pt, eta))
        mask = get_good_node_mask_tensors(pt=pt, particle_id=particle_id, reconstructable=reconstructable,
                                          eta=eta, pt_thld=hp.pt_thld, max_eta=hp.max_eta)
        if hp.sample_pids < 1:  # oc.py:221-225
            mask = mask & (torch.rand_like(beta, dtype=torch.float16) < hp.sample_pids)
        losses, extra = condensation_loss_rg(beta=beta, x=x, particle_id=particle_id, mask=mask, q_min=hp.q_min,
                                             radius_threshold=1.0, max_num_neighbors=hp.max_num_neighbors)
        weights = {"attractive": 1.0, "repulsive": hp.lw_repulsive, "noise": hp.lw_noise, "coward": hp.lw_coward}
        return MultiLossFctReturn(loss_dct=losses, weight_dct=weights, extra_metrics=extra)
