"""Binary on-disk graph format with the destination-sorted plan stored next to the tensors, read
through one pinned staging buffer and one host-to-device copy.

The reference pickles PyG ``Data`` objects and loads one per step (``graph_construction/
graph_builder.py:396-455``, ``utils/loading.py:97-113,219-239``): every step re-derives the
scatter indices and pays a pageable host-to-device copy per tensor.  Here the writer (an offline
converter, like the reference's graph builder) stores

    header   : magic, JSON table {name: dtype, shape, offset} (offsets 4096-byte aligned)
    payload  : x, edge_index, edge_attr, any per-node / per-edge truth tensors, and the plan of
               ``plan.GraphPlan`` -- ``perm`` (stable argsort of ``edge_index[1]``), ``rowptr``,
               ``src_sorted``, ``dst_sorted`` as int32

and the reader brings the whole payload to the device with a single asynchronous copy and adopts the
stored plan (``plan.adopt_plan``), so the first forward over the graph does not sort anything.
``DevicePrefetcher`` / ``GraphLoader`` issue that copy for the next graph on a side stream while the
caller computes on the current one.
"""
from __future__ import annotations

import json
import struct
from pathlib import Path
from typing import Mapping

import numpy as np
import torch
from torch import Tensor

from .plan import GraphPlan, adopt_plan

MAGIC = b"GTBGRAF1"
ALIGN = 4096
_PLAN_KEYS = ("plan.perm", "plan.rowptr", "plan.src_sorted", "plan.dst_sorted")


class GraphData:
    """Attribute bag with the fields the reference's models and losses read from a PyG ``Data``
    (``x``, ``edge_index``, ``edge_attr``, ``y``, ``particle_id``, ``pt``, ...)."""

    def __init__(self, **tensors):
        self.__dict__.update(tensors)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    @property
    def num_nodes(self) -> int:
        return self.x.size(0)

    @property
    def num_edges(self) -> int:
        return self.edge_index.size(1)

    def to(self, device, non_blocking: bool = False) -> "GraphData":
        return GraphData(**{k: (v.to(device, non_blocking=non_blocking) if isinstance(v, Tensor) else v)
                            for k, v in self.__dict__.items() if not k.startswith("_")})


def host_plan(edge_index: np.ndarray, n_nodes: int) -> dict[str, np.ndarray]:
    """The plan of ``gtb_plan_build`` computed by the offline writer: stable sort of the edges by
    destination, CSR row pointers, sorted endpoints (int32)."""
    src, dst = edge_index[0], edge_index[1]
    if src.size and (min(src.min(), dst.min()) < 0 or max(src.max(), dst.max()) >= n_nodes):
        raise IndexError("edge_index contains node indices outside [0, num_nodes)")
    perm = np.argsort(dst, kind="stable").astype(np.int32)
    rowptr = np.zeros(n_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=n_nodes), out=rowptr[1:])
    return {"plan.perm": perm, "plan.rowptr": rowptr, "plan.src_sorted": src[perm].astype(np.int32),
            "plan.dst_sorted": dst[perm].astype(np.int32)}


def write_graph(path, *, x: Tensor, edge_index: Tensor, edge_attr: Tensor, extras: Mapping[str, Tensor] | None = None,
                with_plan: bool = True) -> None:
    arrays: dict[str, np.ndarray] = {
        "x": x.detach().cpu().contiguous().numpy(),
        "edge_index": edge_index.detach().cpu().to(torch.int64).contiguous().numpy(),
        "edge_attr": edge_attr.detach().cpu().contiguous().numpy(),
    }
    for k, v in (extras or {}).items():
        if k in arrays or k.startswith("plan."):
            raise ValueError(f"reserved name {k!r}")
        arrays[k] = v.detach().cpu().contiguous().numpy()
    if with_plan:
        arrays.update(host_plan(arrays["edge_index"], arrays["x"].shape[0]))
    table, off = {}, 0
    for k, a in arrays.items():
        if a.dtype == np.bool_:
            a = arrays[k] = a.view(np.uint8)
            table[k] = {"dtype": "bool", "shape": list(a.shape), "offset": off}
        else:
            table[k] = {"dtype": str(a.dtype), "shape": list(a.shape), "offset": off}
        off = (off + a.nbytes + ALIGN - 1) // ALIGN * ALIGN
    header = json.dumps({"arrays": table, "payload_bytes": off}).encode()
    head_len = (len(MAGIC) + 8 + len(header) + ALIGN - 1) // ALIGN * ALIGN
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", len(header), head_len) + header)
        f.write(b"\0" * (head_len - f.tell()))
        for k, a in arrays.items():
            f.seek(head_len + table[k]["offset"])
            f.write(a.tobytes())
        f.truncate(head_len + off)


_DTYPES = {"float32": torch.float32, "float64": torch.float64, "int64": torch.int64, "int32": torch.int32,
           "uint8": torch.uint8, "int8": torch.int8, "int16": torch.int16, "float16": torch.float16, "bool": torch.bool}


def _read_host(path, pinned: bool):
    """File -> (array table, one host buffer holding the payload)."""
    with open(path, "rb") as f:
        head = f.read(len(MAGIC) + 8)
        if head[:len(MAGIC)] != MAGIC:
            raise ValueError(f"{path}: not a gnn_tracking_b200 graph file")
        n_json, head_len = struct.unpack("<II", head[len(MAGIC):])
        meta = json.loads(f.read(n_json))
        nbytes = int(meta["payload_bytes"])
        host = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=pinned)
        f.seek(head_len)
        got = f.readinto(memoryview(host.numpy())[:nbytes]) if nbytes else 0
        if got != nbytes:
            raise ValueError(f"{path}: truncated payload ({got} of {nbytes} bytes)")
    return meta, host


def _materialize(meta, host: Tensor, device: torch.device) -> GraphData:
    """One asynchronous copy of the payload on the current stream, tensors as views of it, plan adopted."""
    buf = host.to(device, non_blocking=True) if device.type == "cuda" else host
    out: dict[str, Tensor] = {}
    for k, m in meta["arrays"].items():
        dt = _DTYPES[m["dtype"]]
        n_el = int(np.prod(m["shape"])) if m["shape"] else 1
        size = n_el * (1 if dt == torch.bool else torch.empty((), dtype=dt).element_size())
        raw = buf[m["offset"]:m["offset"] + size]
        out[k] = (raw.view(torch.bool) if dt == torch.bool else raw.view(dt)).reshape(m["shape"])
    plan_parts = {k: out.pop(k) for k in _PLAN_KEYS if k in out}
    data = GraphData(**out)
    data._staging = host  # keeps the pinned buffer alive until the asynchronous copy has been consumed
    if len(plan_parts) == len(_PLAN_KEYS) and device.type == "cuda":
        n, e = data.x.size(0), data.edge_index.size(1)
        adopt_plan(data.edge_index, n, GraphPlan(n, e, plan_parts["plan.perm"], plan_parts["plan.rowptr"],
                                                 plan_parts["plan.src_sorted"], plan_parts["plan.dst_sorted"],
                                                 torch.zeros(1, dtype=torch.int32, device=device)))
    else:
        data._plan_arrays = plan_parts
    return data


def read_graph(path, device: torch.device | str = "cuda", *, pinned: bool | None = None) -> GraphData:
    """Load a graph written by ``write_graph``.  The payload is read into one (pinned, when the target
    is a CUDA device) host buffer and copied with a single ``non_blocking`` transfer on the current
    stream; tensors are views of that one device allocation.  The stored plan is adopted."""
    device = torch.device(device)
    pinned = device.type == "cuda" if pinned is None else pinned
    meta, host = _read_host(path, pinned)
    return _materialize(meta, host, device)


_SIDE_STREAMS: dict = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    """One copy stream per device for every prefetcher: the caching allocator keeps a block pool per
    stream, so a fresh stream per loop would start every epoch with cudaMalloc calls."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


class ResultReader:
    """Device-to-host read-back of a step's result on its own stream, so that it runs under the NEXT step's
    kernels instead of holding the compute stream (the mirror image of ``DevicePrefetcher``):

        reader = ResultReader(pinned_host_buffer)
        for data in DevicePrefetcher(...):
            out = model(data)
            reader.read(out["W"])      # returns at once; the copy waits for the kernels that produced W
        reader.wait()                  # host_buffer holds the LAST result

    Copies are ordered among themselves (one stream), so one host buffer is enough as long as the consumer takes
    a result before the next ``read``; ``wait(stream)`` makes a stream (default: the current one) wait instead of
    the host."""

    def __init__(self, host: Tensor, device: torch.device | str = "cuda"):
        self.host = host
        self.device = torch.device(device)
        self._stream = None
        self._done = None

    def read(self, t: Tensor) -> None:
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        t.record_stream(self._stream)  # the allocator must not hand the block out again while the copy reads it
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            self.host[: t.numel()].view(t.shape).copy_(t, non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record(self._stream)

    def wait(self, stream=None, host: bool = False) -> None:
        if self._done is None:
            return
        if host:
            self._done.synchronize()
        else:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self._done)


class DevicePrefetcher:
    """Yields device-resident graphs from an iterable of host-side ones, with the host-to-device
    copy of graph k + 1 issued on a side stream BEFORE graph k is handed to the caller, so the copy
    runs under the caller's kernels for graph k (double buffering; the reference's DataLoader +
    ``.to(device)`` copies in line with the step).  Items are ``GraphData`` objects of (pinned) host
    tensors or the ``(table, staging buffer)`` pairs of ``GraphLoader``."""

    def __init__(self, source, device: torch.device | str = "cuda", *, depth: int = 1):
        self.source = source
        self.device = torch.device(device)
        self.depth = max(1, int(depth))

    def _issue(self, item, stream):
        with torch.cuda.stream(stream):
            if isinstance(item, GraphData):
                data = item.to(self.device, non_blocking=True)
            else:
                data = _materialize(item[0], item[1], self.device)
            ev = torch.cuda.Event()
            ev.record(stream)
        return data, ev

    def __iter__(self):
        from collections import deque
        if self.device.type != "cuda":
            for item in self.source:
                yield item if isinstance(item, GraphData) else _materialize(item[0], item[1], self.device)
            return
        stream = _side_stream(self.device)
        it = iter(self.source)
        inflight: deque = deque()

        def top_up():
            while len(inflight) <= self.depth:
                try:
                    inflight.append(self._issue(next(it), stream))
                except StopIteration:
                    return

        # Back-pressure: kernel launches are asynchronous, so an unthrottled host would enqueue many
        # steps ahead of the GPU; every graph in that backlog pins its own device buffers (they are only
        # reusable once the step that read them has run), and the allocator answers with cudaMalloc calls
        # in the middle of the loop.  The host therefore never gets more than depth + 1 finished-enqueuing
        # steps ahead of the GPU.
        done: deque = deque()
        top_up()
        while inflight:
            data, ev = inflight.popleft()
            top_up()  # the next copy is in flight before the caller launches work on this graph
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for v in data.__dict__.values():
                if isinstance(v, Tensor) and v.is_cuda:
                    v.record_stream(cur)  # allocated on the side stream, used (and freed) on the caller's
            yield data
            del data
            step_done = torch.cuda.Event()
            step_done.record(torch.cuda.current_stream(self.device))  # the caller's work on this graph is enqueued
            done.append(step_done)
            if len(done) > self.depth + 1:
                done.popleft().synchronize()


def shard_paths(paths, rank: int, world_size: int) -> list:
    """The files of one rank when ``world_size`` processes walk the same list, one graph per step each (the
    reference trains through Lightning's DDP strategy, whose ``DistributedSampler(shuffle=False)`` this follows):
    the list is padded by wrapping around to a multiple of ``world_size`` so that every rank takes the same number
    of steps (the gradient all-reduce needs them in lockstep), then rank r takes entries r, r + W, r + 2W, ..."""
    paths = list(paths)
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside a world of {world_size}")
    if world_size == 1 or not paths:
        return paths
    total = -(-len(paths) // world_size) * world_size
    padded = paths + [paths[i % len(paths)] for i in range(total - len(paths))]
    return padded[rank::world_size]


def gpu_local_cpus(device: torch.device | str | int) -> set[int]:
    """CPUs on the NUMA node the GPU hangs off (NVML's ideal affinity), restricted to the ones this process may
    use; empty if NVML cannot tell."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        index = torch.device(device).index if not isinstance(device, int) else device
        if index is None:
            index = torch.cuda.current_device()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:  # NVML numbers the physical devices
            entry = visible.split(",")[index].strip()
            handle = pynvml.nvmlDeviceGetHandleByUUID(entry) if entry.startswith("GPU-") else \
                pynvml.nvmlDeviceGetHandleByIndex(int(entry))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
    except Exception:  # noqa: BLE001 - no NVML, no affinity information, a container without the device node
        return set()
    cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
    return cpus & os.sched_getaffinity(0)


def bind_thread_near_gpu(device: torch.device | str | int) -> bool:
    """Pin the CALLING thread to the CPUs next to the GPU, so that the pinned staging buffers it allocates and
    fills are first touched on that NUMA node and the host-to-device copies do not cross the socket link.
    Returns False (and changes nothing) when the topology is unknown."""
    import os
    cpus = gpu_local_cpus(device)
    if not cpus:
        return False
    os.sched_setaffinity(0, cpus)  # tid 0 = the calling thread on Linux
    return True


class GraphLoader:
    """Iterates over graph files one graph per step (the reference trains with ``batch_size=1``,
    ``utils/loading.py:235``): a reader thread fills pinned staging buffers ``prefetch`` files ahead
    and a ``DevicePrefetcher`` keeps the copy of the next graph in flight under the current step.

    One process per GPU: every rank builds the loader over the SAME file list and walks its own share
    (``shard_paths``; ``rank`` / ``world_size`` default to the initialised process group, else one process).
    ``numa_local=True`` runs the reader thread on the CPUs next to this rank's GPU (``bind_thread_near_gpu``)."""

    def __init__(self, paths, device: torch.device | str = "cuda", *, prefetch: int = 2, pinned: bool | None = None,
                 rank: int | None = None, world_size: int | None = None, numa_local: bool = False):
        if rank is None or world_size is None:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            rank = (dist.get_rank() if on else 0) if rank is None else rank
            world_size = (dist.get_world_size() if on else 1) if world_size is None else world_size
        self.paths = [Path(p) for p in shard_paths(paths, rank, world_size)]
        self.rank, self.world_size = rank, world_size
        self.device = torch.device(device)
        self.pinned = self.device.type == "cuda" if pinned is None else pinned
        self.prefetch = max(1, int(prefetch))
        self.numa_local = bool(numa_local) and self.device.type == "cuda"

    def __len__(self) -> int:
        return len(self.paths)

    def _host_items(self):
        import queue
        import threading

        q: queue.Queue = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def reader():
            try:
                if self.numa_local:
                    bind_thread_near_gpu(self.device)
                for p in self.paths:
                    if stop.is_set():
                        return
                    q.put(_read_host(p, self.pinned))
                q.put(None)
            except BaseException as exc:  # noqa: BLE001 - re-raised in the consumer
                q.put(exc)

        th = threading.Thread(target=reader, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while th.is_alive():  # unblock a reader waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    th.join(timeout=0.05)

    def __iter__(self):
        return iter(DevicePrefetcher(self._host_items(), self.device))
