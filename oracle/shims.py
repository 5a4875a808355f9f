"""TEST INFRASTRUCTURE ONLY -- stub modules that let the reference's own source
files (``/root/reference/src/gnn_tracking``) import in a container that has no
``torch_geometric`` / ``pytorch_lightning`` / ``torch_cluster`` / ``colorlog`` /
``torchmetrics``.

Nothing in the product package imports this file.  It is used by
``tests/golden/make_golden.py`` (run in the authoring container, where
``/root/reference`` exists) to generate the committed golden vectors, and by the
CPU tests that validate ``oracle/*.py`` against the real reference when the
reference is present.

The third-party semantics restated here:

* ``torch_geometric.nn.MessagePassing.propagate`` for a dense ``edge_index``,
  ``aggr="add"``, ``flow="source_to_target"`` (torch_geometric >= 2.3, pinned only
  as ``>=2.3.0`` at reference ``setup.cfg:43``): ``*_j`` args are
  ``index_select(0, edge_index[0])``, ``*_i`` args ``index_select(0, edge_index[1])``,
  the message is summed with ``new_zeros(N, F).scatter_add_(0, edge_index[1], msg)``.
* ``torch_cluster.radius_graph`` (unpinned, reference ``environments/default.yml:15``):
  brute force, strict ``<`` on the distance, no self loops, rows ``[neighbour, centre]``.
* ``torch_geometric.data.Data``: attribute bag with ``edge_subgraph`` / ``subgraph``.
"""
from __future__ import annotations

import copy
import inspect
import logging
import sys
import types

import torch
from torch import nn


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # deepcopy / hasattr rely on AttributeError
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class HyperparametersMixin:
    """Restatement of lightning's mixin: ``save_hyperparameters`` captures the
    caller frame's ctor locals (or merges a dict argument) into ``self.hparams``."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)

    @property
    def hparams(self):
        if "_hparams" not in self.__dict__:
            self.__dict__["_hparams"] = _AttrDict()
        return self.__dict__["_hparams"]

    def save_hyperparameters(self, *args, ignore=None, frame=None, logger=True):
        if ignore is None:
            ignore = []
        if isinstance(ignore, str):
            ignore = [ignore]
        if args and isinstance(args[0], dict):
            self.hparams.update(args[0])
            return
        frame = frame or inspect.currentframe().f_back
        loc = frame.f_locals
        code = frame.f_code
        names = code.co_varnames[: code.co_argcount + code.co_kwonlyargcount]
        out = {}
        for n in names:
            if n in ("self",) or n in ignore:
                continue
            if n in loc:
                out[n] = loc[n]
        # **kwargs of the ctor are flattened
        if code.co_flags & inspect.CO_VARKEYWORDS:
            kwname = code.co_varnames[
                code.co_argcount
                + code.co_kwonlyargcount
                + (1 if code.co_flags & inspect.CO_VARARGS else 0)
            ]
            for k, v in (loc.get(kwname) or {}).items():
                if k not in ignore:
                    out[k] = v
        self.hparams.update(out)


class MessagePassing(nn.Module):
    """PyG ``MessagePassing`` restated for Tensor ``edge_index`` and sum aggregation."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kw):
        super().__init__()
        assert aggr in ("add", "sum"), "oracle shim only restates sum aggregation"
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        n = kwargs["x"].size(0)
        margs = {}
        for a in list(inspect.signature(self.message).parameters):
            if a.endswith("_j"):
                margs[a] = kwargs[a[:-2]].index_select(0, edge_index[0])
            elif a.endswith("_i"):
                margs[a] = kwargs[a[:-2]].index_select(0, edge_index[1])
            else:
                margs[a] = kwargs[a]
        msg = self.message(**margs)
        idx = edge_index[1].view(-1, 1).expand_as(msg)
        aggr = msg.new_zeros(n, msg.size(1)).scatter_add_(0, idx, msg)
        uparams = list(inspect.signature(self.update).parameters)[1:]
        return self.update(aggr, **{a: kwargs[a] for a in uparams})


class Data:
    """Attribute bag standing in for ``torch_geometric.data.Data``."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    # -- pickle support for the bundled tests/test_data/graphs/test_graph.pt -----
    def __setstate__(self, state):
        self.__dict__.update(state)
        store = self.__dict__.get("_store")
        if store is not None:
            mapping = getattr(store, "_mapping", None) or store.__dict__.get("_mapping", {})
            for k, v in mapping.items():
                self.__dict__[k] = v

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def __contains__(self, k):
        return k in self.__dict__

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def _clone_with(self, fn):
        out = copy.copy(self)
        out.__dict__ = dict(self.__dict__)
        for k in self.keys():
            out.__dict__[k] = fn(k, getattr(self, k))
        return out

    @staticmethod
    def _cat_dim(key):
        return -1 if "index" in key else 0

    def _is_edge_attr(self, k, v):
        """PyG ``BaseStorage.is_edge_attr``: a tensor whose cat-dim equals num_edges
        (ties with num_nodes broken by 'edge' in the key)."""
        if not torch.is_tensor(v) or v.dim() == 0:
            return False
        if v.shape[self._cat_dim(k)] != self.num_edges:
            return False
        if self.num_nodes != self.num_edges:
            return True
        return "edge" in k

    def _is_node_attr(self, k, v):
        if not torch.is_tensor(v) or v.dim() == 0:
            return False
        if v.shape[self._cat_dim(k)] != self.num_nodes:
            return False
        if self.num_nodes != self.num_edges:
            return True
        return "edge" not in k

    def edge_subgraph(self, subset):
        """PyG ``Data.edge_subgraph``: keeps all nodes; filters ``edge_index``
        columns and every edge-level attribute."""

        def f(k, v):
            if k == "edge_index":
                return v[:, subset]
            if self._is_edge_attr(k, v):
                return v[subset] if self._cat_dim(k) == 0 else v[..., subset]
            return v

        return self._clone_with(f)

    def subgraph(self, subset):
        """PyG ``Data.subgraph``: node-induced subgraph with relabelled nodes."""
        n = self.num_nodes
        if subset.dtype == torch.bool:
            node_mask = subset
            subset = node_mask.nonzero().view(-1)
        else:
            node_mask = torch.zeros(n, dtype=torch.bool)
            node_mask[subset] = True
        relabel = torch.full((n,), -1, dtype=torch.long)
        relabel[subset] = torch.arange(subset.numel())
        ei = self.edge_index
        edge_mask = node_mask[ei[0]] & node_mask[ei[1]]

        def f(k, v):
            if k == "edge_index":
                return relabel[v[:, edge_mask]]
            if self._is_node_attr(k, v):
                return v[subset]
            if self._is_edge_attr(k, v):
                return v[edge_mask]
            return v

        return self._clone_with(f)


def index_to_mask(index, size=None):
    size = int(index.max()) + 1 if size is None else size
    mask = index.new_zeros(size, dtype=torch.bool)
    mask[index] = True
    return mask


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", num_workers=1):
    """Brute-force ``torch_cluster.radius_graph``: row 0 = neighbour, row 1 = centre."""
    d = torch.cdist(x, x)
    adj = d < r
    if batch is not None:
        adj &= batch.view(-1, 1) == batch.view(1, -1)
    if not loop:
        adj.fill_diagonal_(False)
    centre, neigh = adj.nonzero(as_tuple=True)
    # torch_cluster caps the number of neighbours per centre; order is
    # implementation defined, here: ascending neighbour index.
    if max_num_neighbors is not None:
        rank = torch.zeros_like(centre)
        if centre.numel():
            start = torch.ones_like(centre, dtype=torch.bool)
            start[1:] = centre[1:] != centre[:-1]
            seg_start = torch.where(start)[0]
            seg_id = torch.cumsum(start.long(), 0) - 1
            rank = torch.arange(centre.numel()) - seg_start[seg_id]
        keep = rank < max_num_neighbors
        centre, neigh = centre[keep], neigh[keep]
    return torch.stack([neigh, centre])


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Inject the stub modules (idempotent; never overrides a real install)."""
    if "pytorch_lightning" not in sys.modules:
        try:
            import pytorch_lightning  # noqa: F401
        except Exception:
            pl = _mod("pytorch_lightning", LightningModule=nn.Module)
            _mod("pytorch_lightning.callbacks", ProgressBar=object)
            pl.callbacks = sys.modules["pytorch_lightning.callbacks"]
            core = _mod("pytorch_lightning.core")
            mix = _mod("pytorch_lightning.core.mixins")
            hp = _mod("pytorch_lightning.core.mixins.hparams_mixin", HyperparametersMixin=HyperparametersMixin)
            pl.core, core.mixins, mix.hparams_mixin = core, mix, hp
    if "torch_geometric" not in sys.modules:
        try:
            import torch_geometric  # noqa: F401
        except Exception:
            tg = _mod("torch_geometric")
            tg.nn = _mod("torch_geometric.nn", MessagePassing=MessagePassing)
            tg.data = _mod("torch_geometric.data", Data=Data)
            dd = _mod("torch_geometric.data.data", Data=Data, DataTensorAttr=Data, DataEdgeAttr=Data)
            st = _mod("torch_geometric.data.storage", GlobalStorage=Data, BaseStorage=Data, NodeStorage=Data, EdgeStorage=Data)
            tg.data.data, tg.data.storage = dd, st
            tg.utils = _mod("torch_geometric.utils", index_to_mask=index_to_mask)
            tg.nn.conv = _mod("torch_geometric.nn.conv", MessagePassing=MessagePassing)
            tg.nn.__path__ = []  # mark as package so "torch_geometric.nn.conv" resolves
            from typing import Optional, Tuple
            OptT = Optional[torch.Tensor]
            tg.typing = _mod("torch_geometric.typing", OptTensor=OptT, PairTensor=Tuple[torch.Tensor, torch.Tensor], PairOptTensor=Tuple[OptT, OptT])
            try:
                torch.serialization.add_safe_globals([Data])
            except Exception:
                pass
    if "torch_cluster" not in sys.modules:
        try:
            import torch_cluster  # noqa: F401
        except Exception:
            _mod("torch_cluster", radius_graph=radius_graph, knn=None, knn_graph=None)
    if "colorlog" not in sys.modules:
        try:
            import colorlog  # noqa: F401
        except Exception:
            class _Fmt(logging.Formatter):
                def __init__(self, fmt=None, *a, **k):
                    k.pop("log_colors", None)
                    k.pop("reset", None)
                    k.pop("secondary_log_colors", None)
                    super().__init__((fmt or "%(message)s").replace("%(log_color)s", "").replace("%(reset)s", ""), *[x for x in a if isinstance(x, str)])
            _mod("colorlog", getLogger=logging.getLogger, StreamHandler=logging.StreamHandler, ColoredFormatter=_Fmt)
    if "torchmetrics" not in sys.modules:
        try:
            import torchmetrics  # noqa: F401
        except Exception:
            _mod("torchmetrics", Metric=nn.Module)
