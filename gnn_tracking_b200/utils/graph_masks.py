"""Hit-of-interest mask (reference utils/graph_masks.py:19-35)."""
from __future__ import annotations

from torch import Tensor


def get_good_node_mask_tensors(*, pt: Tensor, particle_id: Tensor, reconstructable: Tensor, eta: Tensor,
                               pt_thld: float = 0.9, max_eta: float = 4.0) -> Tensor:
    return (pt > pt_thld) & (particle_id > 0) & (reconstructable > 0) & (eta.abs() < max_eta)


def get_good_node_mask(data, *, pt_thld: float = 0.9, max_eta: float = 4.0) -> Tensor:
    return get_good_node_mask_tensors(pt=data.pt, particle_id=data.particle_id,
                                      reconstructable=data.reconstructable, eta=data.eta,
                                      pt_thld=pt_thld, max_eta=max_eta)
