"""DBSCAN for the cluster hyper-parameter scan, on the GPU (reference
postprocessing/fastrescanner.py:6-66; caller postprocessing/dbscanscanner.py:146-187).

The reference copies the latent coordinates ``H`` to the host, lets sklearn build one radius-neighbour
graph at ``max_eps`` and re-filters it per trial before calling sklearn's Cython ``dbscan_inner``.
Here every trial is one ``gtb_dbscan_f32`` call on the device-resident ``H`` (three brute-force
passes over shared-memory tiles: core flags, union-find of the core samples, labels); the labels
are identical to sklearn's, including their numbering."""
from __future__ import annotations

import os

import torch
from torch import Tensor

from .. import ops
from .._lib import check, lib


def dbscan(x: Tensor, eps: float = 1.0, min_samples: int = 1, *, method: str | None = None) -> Tensor:
    """``sklearn.cluster.DBSCAN(eps, min_samples).fit_predict(x)`` as an int64 tensor on ``x``'s
    device (-1 = noise).  ``method``: "grid" (default: uniform cell list, ``gtb_dbscan_grid_f32``) or
    "brute" (all pairs, ``gtb_dbscan_f32``); both give the same labels (``GTB_DBSCAN`` overrides the default)."""
    dev = ops.require_cuda(x)
    x = x.detach().to(torch.float32).contiguous()
    n, d = x.shape
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    core = torch.empty(n, dtype=torch.uint8, device=dev)
    parent = torch.empty(n, dtype=torch.int32, device=dev)
    root = torch.empty(n, dtype=torch.int32, device=dev)
    method = method or os.environ.get("GTB_DBSCAN", "grid")
    if method == "grid":
        ws_bytes = lib().gtb_dbscan_grid_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().gtb_dbscan_grid_f32(x.data_ptr(), d, n, float(eps), int(min_samples), core.data_ptr(), parent.data_ptr(),
                                        root.data_ptr(), ws.data_ptr(), ws_bytes, ops.stream_ptr(dev)))
        ops._count(10)
    elif method == "brute":
        check(lib().gtb_dbscan_f32(x.data_ptr(), d, n, float(eps), int(min_samples), core.data_ptr(), parent.data_ptr(),
                                   root.data_ptr(), ops.stream_ptr(dev)))
        ops._count(3)
    else:
        raise ValueError(f"unknown DBSCAN method {method!r}")
    # number the clusters in increasing order of their lowest core index (dbscan_inner's order)
    is_root = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    clustered = root >= 0
    is_root[root[clustered].long()] = 1
    rank = torch.cumsum(is_root, 0) - 1
    labels = torch.full((n,), -1, dtype=torch.int64, device=dev)
    labels[clustered] = rank[root[clustered].long()]
    return labels


class DBSCANFastRescan:
    """Interface of the reference class (fastrescanner.py:6-66): ``cluster(eps, min_pts)`` for many
    trials over the same points.  ``x`` is a CUDA tensor (labels come back as a CUDA int64 tensor) or,
    as the reference's scanner passes it (dbscanscanner.py:160-165), a numpy array: it is moved to
    the device once and every trial returns a numpy label array like the reference.  ``max_eps`` and
    ``n_jobs`` are accepted for compatibility (no neighbour graph is cached: a trial is a few
    milliseconds)."""

    def __init__(self, x, max_eps: float = 1.0, *, n_jobs: int | None = None):
        self._numpy = not isinstance(x, Tensor)
        if self._numpy:
            import numpy as np
            x = torch.from_numpy(np.ascontiguousarray(x)).to("cuda")
        ops.require_cuda(x)
        self.x = x.detach().to(torch.float32).contiguous()
        self._max_eps = max_eps

    def cluster(self, eps: float = 1.0, min_pts: int = 1):
        labels = dbscan(self.x, eps, min_pts)
        return labels.cpu().numpy() if self._numpy else labels
