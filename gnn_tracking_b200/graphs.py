"""The whole forward of a graph-shaped model as ONE CUDA graph launch.

An edge-classifier forward is ~25 kernel launches of 5 - 300 us each (plan build, encoders, two launches per
Interaction-Network layer, W head): the host needs ~0.8 ms to enqueue what the GPU runs in ~1.3 ms, so any
hiccup of the host -- another process on its cores, a profiler thread, eight ranks sharing one socket -- shows up
in the step time.  For a fixed graph SHAPE (N nodes, E edges; the values change every step) the launch sequence is
static: ``CapturedForward`` records it once into a ``torch.cuda.CUDAGraph`` over static input buffers and replays it
per step -- plan build (destination sort of the NEW edge_index) included, nothing is cached between steps.

    fwd = CapturedForward(model, x, edge_index, edge_attr)       # warms up, captures
    out = fwd(x2, edge_index2, edge_attr2)                       # three device copies + one graph launch
    out["W"] ...                                                 # static output buffers, valid until the next call

No gradients (inference / evaluation / the throughput benchmark); other shapes re-capture (``fits``)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops
from .plan import clear_plan_cache


class CapturedForward:
    def __init__(self, model, x: Tensor, edge_index: Tensor, edge_attr: Tensor, *, halo=None, warmup: int = 2):
        dev = ops.require_cuda(x, edge_index, edge_attr)
        self.model, self.halo = model, halo
        self.x, self.edge_index, self.edge_attr = x.clone(), edge_index.clone(), edge_attr.clone()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):  # packs the weights, sets the kernel attributes, sizes the allocator pools
                clear_plan_cache()
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        _drop_scratch(model)
        l0 = ops.launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            clear_plan_cache()
            self.out = self._run()
        self.launches = ops.launch_count() - l0  # kernels of the library inside one replay
        self._weights = self._weight_key()
        clear_plan_cache()       # the captured plan lives in the graph's memory pool: not for eager callers
        _drop_scratch(model)

    def _run(self):
        kw = {} if self.halo is None else {"halo": self.halo}
        return self.model.forward_tensors(self.x, self.edge_index, self.edge_attr, **kw)

    def fits(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> bool:
        return (x.shape == self.x.shape and edge_index.shape == self.edge_index.shape
                and edge_attr.shape == self.edge_attr.shape)

    def load(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> None:
        """Copies a new graph of the captured shape into the static input buffers (device or pinned host tensors)."""
        if not self.fits(x, edge_index, edge_attr):
            raise ValueError("graph shape differs from the captured one: capture again")
        self.x.copy_(x, non_blocking=True)
        self.edge_index.copy_(edge_index, non_blocking=True)
        self.edge_attr.copy_(edge_attr, non_blocking=True)

    def _weight_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def replay(self):
        if self._weight_key() != self._weights:  # the graph holds the PACKED weights of the capture
            raise RuntimeError("the model's weights changed since the capture (optimizer step, load_state_dict, .to()): "
                               "capture again")
        ops._count(self.launches)
        self.graph.replay()
        return self.out

    def __call__(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor):
        self.load(x, edge_index, edge_attr)
        return self.replay()


def _drop_scratch(model) -> None:
    """Scratch the stacks keep between forwards (the zeroed aggregate) must not cross the eager / captured boundary."""
    for m in model.modules():
        m.__dict__.pop("_aggr_buf", None)
