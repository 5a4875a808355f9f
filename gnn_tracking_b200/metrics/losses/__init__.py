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
