"""Return-type contract of the losses on the path (reference
metrics/losses/__init__.py:13-54)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import torch
from torch import Tensor


@dataclass(kw_only=True)
class MultiLossFctReturn:
    """Split losses + weights; ``.loss`` is the weighted sum consumed by the training
    modules (reference training/tc.py:64-70)."""
    loss_dct: dict[str, Tensor]
    weight_dct: dict[str, Tensor] | dict[str, float]
    extra_metrics: dict[str, Any] = field(default_factory=dict)

    def __post_init__(self) -> None:
        assert self.loss_dct.keys() == self.weight_dct.keys()

    @property
    def weighted_losses(self) -> dict[str, Tensor]:
        return {k: v * self.weight_dct[k] for k, v in self.loss_dct.items()}

    @property
    def loss(self) -> Tensor:
        total = sum(self.weighted_losses.values())
        assert isinstance(total, torch.Tensor)
        return total


class MultiLossFct(torch.nn.Module):
    def forward(self, *args: Any, **kwargs: Any) -> MultiLossFctReturn: ...


class DummyMultiLoss(MultiLossFct):
    """``sum(x)`` as the only loss term: times a training loop without a real loss (reference
    metrics/losses/__init__.py:44-54)."""

    def forward(self, x: Tensor, **kwargs: Any) -> MultiLossFctReturn:
        return MultiLossFctReturn(loss_dct={"dummy": torch.sum(x)}, weight_dct={"dummy": 1.0})


class LossClones(torch.nn.Module):
    def __init__(self, loss: torch.nn.Module, prefixes=("w", "y")) -> None:
        """Evaluates ``loss`` once per suffixed model output (reference
        metrics/losses/__init__.py:57-131): with outputs ``w_0, w_1, ...`` (and truths ``y_0, ...``)
        the wrapped loss sees them as ``w`` / ``y`` and the result is ``{"0": loss_0, "1": ...}``.  The
        suffixes come from the first prefix; an un-suffixed output named like a prefix is dropped."""
        super().__init__()
        self._loss = loss
        self._prefixes = tuple(prefixes)

    def forward(self, **kwargs) -> dict[str, Tensor]:
        outputs = {k: v for k, v in kwargs.items() if k not in self._prefixes}
        lead = self._prefixes[0] + "_"
        out = {}
        for suffix in sorted(k[len(lead):] for k in outputs if k.startswith(lead)):
            rename = {f"{p}_{suffix}": p for p in self._prefixes}
            out[suffix] = self._loss(**{rename.get(k, k): v for k, v in outputs.items()})
        return out
