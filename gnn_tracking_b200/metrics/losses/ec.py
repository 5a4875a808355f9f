"""Edge-classification losses behind the reference interface (reference
metrics/losses/ec.py:95-178): one fused reduction kernel per call
(``gtb_ec_loss_f32``): label falsification by ``pt[edge_index[0]]``, the per-edge
BCE / focal term and the sum, finished as ``sum / n`` on the device."""
from __future__ import annotations

import math

import torch
from torch import Tensor

from ... import ops
from ..._hparams import HyperparametersMixin
from ..._lib import check, lib

_BCE, _FOCAL, _HAUGHTY = 0, 1, 2


class _ECLossFn(torch.autograd.Function):
    """Mean EC loss with its analytic gradient w.r.t. the edge weights (``gtb_ec_loss_grad_f32``)."""

    @staticmethod
    def forward(ctx, w, yk, kind, src, ptf, pt_thld, mode, alpha, gamma, pos_weight):
        dev = w.device
        out = torch.zeros(2, dtype=torch.float64, device=dev)
        check(lib().gtb_ec_loss_f32(w.data_ptr(), yk.data_ptr(), kind, w.numel(), None if src is None else src.data_ptr(),
                                    None if ptf is None else ptf.data_ptr(), float(pt_thld), mode, float(alpha), float(gamma),
                                    float(pos_weight), out.data_ptr(), ops.stream_ptr(dev)))
        ops._count(1)
        ctx.save_for_backward(w, yk, *([src, ptf] if src is not None else []))
        ctx.cfg = (kind, float(pt_thld), mode, float(alpha), float(gamma), float(pos_weight), src is not None)
        return (out[0] / out[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, grad):
        kind, pt_thld, mode, alpha, gamma, pos_weight, has_pt = ctx.cfg
        saved = ctx.saved_tensors
        w, yk = saved[0], saved[1]
        src, ptf = (saved[2], saved[3]) if has_pt else (None, None)
        scale = (grad.detach().to(torch.float32) / max(w.numel(), 1)).reshape(1).contiguous()
        dw = torch.empty_like(w)
        check(lib().gtb_ec_loss_grad_f32(w.data_ptr(), yk.data_ptr(), kind, w.numel(), None if src is None else src.data_ptr(),
                                         None if ptf is None else ptf.data_ptr(), pt_thld, mode, alpha, gamma, pos_weight,
                                         scale.data_ptr(), dw.data_ptr(), ops.stream_ptr(w.device)))
        ops._count(1)
        return (dw,) + (None,) * 9


def _ec_loss(w: Tensor, y: Tensor, *, mode: int, edge_index: Tensor | None, pt: Tensor | None, pt_thld: float,
             alpha: float = 0.25, gamma: float = 2.0, pos_weight: float = 1.0) -> Tensor:
    ops.require_cuda(w, y)
    shape_w = w.shape
    w = w.reshape(-1)
    if w.dtype != torch.float32:
        raise TypeError("w must be float32")
    w = w.contiguous()
    y = y.reshape(-1)
    if y.dtype in (torch.bool, torch.uint8):
        yk, kind = y.contiguous().view(torch.uint8), 1
    else:
        yk, kind = y.to(torch.float32).contiguous(), 0
    src = ptf = None
    if not math.isclose(pt_thld, 0.0):  # falsify_low_pt_edges ec.py:71-92
        assert edge_index is not None and pt is not None
        src = edge_index[0].contiguous()
        ptf = pt.to(torch.float32).contiguous()
    del shape_w
    return _ECLossFn.apply(w, yk, kind, src, ptf, pt_thld, mode, alpha, gamma, pos_weight)


class FalsifyLowPtEdgeWeightLoss(torch.nn.Module, HyperparametersMixin):
    def __init__(self, *, pt_thld: float = 0.0):
        super().__init__()
        self.save_hyperparameters()


class EdgeWeightBCELoss(FalsifyLowPtEdgeWeightLoss):
    """``mean(binary_cross_entropy(w, falsified y))`` (ec.py:116-121)."""

    def forward(self, *, w: Tensor, y: Tensor, edge_index: Tensor | None = None, pt: Tensor | None = None,
                **kwargs) -> Tensor:
        return _ec_loss(w, y, mode=_BCE, edge_index=edge_index, pt=pt, pt_thld=self.hparams.pt_thld)


class EdgeWeightFocalLoss(FalsifyLowPtEdgeWeightLoss):
    def __init__(self, *, alpha=0.25, gamma=2.0, pos_weight=None, **kwargs):
        """Focal loss (ec.py:124-150, core :12-29)."""
        super().__init__(**kwargs)
        self.save_hyperparameters()

    def forward(self, *, w: Tensor, y: Tensor, edge_index: Tensor | None = None, pt: Tensor | None = None,
                **kwargs) -> Tensor:
        pw = self.hparams.pos_weight
        pw = 1.0 if pw is None else float(torch.as_tensor(pw).reshape(-1)[0])
        return _ec_loss(w, y, mode=_FOCAL, edge_index=edge_index, pt=pt, pt_thld=self.hparams.pt_thld,
                        alpha=self.hparams.alpha, gamma=self.hparams.gamma, pos_weight=pw)


class HaughtyFocalLoss(torch.nn.Module, HyperparametersMixin):
    def __init__(self, *, alpha: float = 0.25, gamma: float = 2.0, pt_thld=0.0):
        """Focal loss whose positive weight is the pt-falsified label while the raw label
        stays the target (ec.py:153-178)."""
        super().__init__()
        self.save_hyperparameters()
        self._alpha, self._gamma, self._pt_thld = alpha, gamma, pt_thld

    def forward(self, *, w: Tensor, y: Tensor, edge_index: Tensor, pt: Tensor, **kwargs) -> Tensor:
        return _ec_loss(w, y, mode=_HAUGHTY, edge_index=edge_index, pt=pt, pt_thld=self._pt_thld,
                        alpha=self._alpha, gamma=self._gamma)
