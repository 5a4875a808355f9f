"""Host-side wrappers of the C ABI: tensors in, tensors out.  PyTorch is used for
device memory and streams only; all arithmetic happens in ``libgtb200.so``."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Sequence

import torch
from torch import Tensor

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID_AFFINE, IMPL_AUTO, IMPL_FFMA, IMPL_TCGEN05, SRC_PROJECTED,
                   SRC_SORTED, GtbError, MlpDesc, Src, check, lib)

__all__ = ["ACT_NONE", "ACT_RELU", "ACT_SIGMOID_AFFINE", "IMPL_AUTO", "IMPL_FFMA", "IMPL_TCGEN05", "Block",
           "PackedMLP", "fused_mlp", "pack_linears", "require_cuda", "default_impl", "launch_count", "rows_gather",
           "tc_slots", "in_edge_bf16", "pack_in_edge_bf16", "in_node_fused", "edge_encoder"]

_LAUNCHES = 0  # kernels launched through the C ABI by this process (bench.py reports it)


def launch_count() -> int:
    return _LAUNCHES


def _count(n: int) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def require_cuda(*tensors: Tensor) -> torch.device:
    """The product path is CUDA sm_100a only: there is no CPU / eager fallback."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "gnn_tracking_b200 runs on a CUDA sm_100a (B200) device only; got a tensor on "
                f"{t.device}.  There is no CPU fallback (use the reference package on CPU).")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    if dev is None:
        raise RuntimeError("no tensor given")
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    return dev


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class on_device:
    """Makes ``dev`` the current CUDA device around a library call when it is not already (tensors on
    cuda:1 while cuda:0 is current): kernel attributes and launches are per device.  A no-op -- one
    integer compare -- in the usual one-process-per-GPU set-up."""

    __slots__ = ("_guard",)

    def __init__(self, dev: torch.device):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self._guard = torch.cuda.device(idx) if idx != torch.cuda.current_device() else None

    def __enter__(self):
        if self._guard is not None:
            self._guard.__enter__()

    def __exit__(self, *exc):
        if self._guard is not None:
            self._guard.__exit__(*exc)


def default_impl() -> int:
    """GTB_IMPL_AUTO unless ``GTB_IMPL`` = ffma | tcgen05 forces one."""
    v = os.environ.get("GTB_IMPL", "auto").lower()
    return {"auto": IMPL_AUTO, "ffma": IMPL_FFMA, "tcgen05": IMPL_TCGEN05}[v]


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"gtb200 kernels are fp32; got {t.dtype}")
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.stride(-1) != 1 and t.size(-1) > 1:
        t = t.contiguous()
    return t


@dataclass
class Block:
    """One column block of a concatenated MLP input: ``act(tensor[index])``.

    ``projected``: ``tensor`` already holds the block multiplied by its columns of the first
    Linear (a ``[*, N0]`` table); its gathered rows are added behind the first Linear instead
    of being concatenated in front of it (GTB_SRC_PROJECTED in include/gtb200.h).
    ``sorted_index``: hint that ``index`` is non-decreasing (the plan's ``dst_sorted``)."""
    tensor: Tensor
    index: Tensor | None = None  # int32 row index per output row
    relu: bool = False
    projected: bool = False
    sorted_index: bool = False
    unique_index: bool = False   # hint: ``index`` is a permutation of the table's rows (the plan's ``perm``)
    # node-partitioned graphs: maps the per-node table of the OWNED rows (the block itself, or its
    # pre-projected table) to owned + halo rows (partition.HaloExchange.extend); ``index`` then
    # addresses the extended table
    extend: object = None

    @property
    def flags(self) -> int:
        return (SRC_PROJECTED if self.projected else 0) | (SRC_SORTED if self.sorted_index else 0)


@dataclass
class PackedMLP:
    buf: Tensor
    dims: tuple[int, ...]
    impl: int
    block_widths: tuple[int, ...] = ()

    @property
    def n_layers(self) -> int:
        return len(self.dims) - 1


def _i32arr(vals: Sequence[int]):
    return (C.c_int32 * len(vals))(*vals)


def resolve_impl(dims: Sequence[int], impl: int, block_widths: Sequence[int] | None = None) -> int:
    """GTB_IMPL_AUTO -> the tcgen05 tiles when they support the widths, else the FFMA tiles."""
    if impl != IMPL_AUTO:
        return impl
    bw = list(block_widths) if block_widths else [dims[0]]
    ok = lib().gtb_mlp_packed_bytes(len(dims) - 1, _i32arr(dims), len(bw), _i32arr(bw), IMPL_TCGEN05) > 0
    return IMPL_TCGEN05 if ok else IMPL_FFMA


def tc_slots(dims: Sequence[int], block_widths: Sequence[int] | None = None) -> int:
    """Staging slots the tcgen05 tiles would have beside these weights (0 = unsupported)."""
    bw = list(block_widths) if block_widths else [dims[0]]
    return lib().gtb_mlp_tc_slots(len(dims) - 1, _i32arr(dims), len(bw), _i32arr(bw))


def pack_linears(weights: Sequence[Tensor], biases: Sequence[Tensor | None], impl: int = IMPL_AUTO,
                 block_widths: Sequence[int] | None = None) -> PackedMLP:
    """Repack <= 3 ``nn.Linear`` layers ([out, in] weights) for the fused tiles.
    ``block_widths``: widths of the concatenated source blocks the MLP is called with."""
    dev = require_cuda(*weights)
    n = len(weights)
    if not 1 <= n <= _lib.GTB_MAX_LAYERS:
        raise ValueError(f"a fused MLP call takes 1..{_lib.GTB_MAX_LAYERS} Linear layers, got {n}")
    dims = [weights[0].size(1)] + [w.size(0) for w in weights]
    for i, w in enumerate(weights):
        if w.size(1) != dims[i]:
            raise ValueError("Linear layers do not chain")
    bw = list(block_widths) if block_widths else [dims[0]]
    if sum(bw) != dims[0]:
        raise ValueError(f"block widths {bw} do not sum to the input width {dims[0]}")
    impl = resolve_impl(dims, impl, bw)
    dims_c = _i32arr(dims)
    bw_c = _i32arr(bw)
    nbytes = lib().gtb_mlp_packed_bytes(n, dims_c, len(bw), bw_c, impl)
    if nbytes == 0:
        raise GtbError(-2, f"MLP widths {dims} (blocks {bw}) are not supported by impl {impl}")
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ws = [w.detach().to(torch.float32).contiguous() for w in weights]
    bs = [None if b is None else b.detach().to(torch.float32).contiguous() for b in biases]
    wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
    bp = (C.c_void_p * n)(*[(b.data_ptr() if b is not None else None) for b in bs])
    with on_device(dev):
        check(lib().gtb_mlp_pack(n, dims_c, len(bw), bw_c, wp, bp, impl, buf.data_ptr(), stream_ptr(dev)))
    _count(n if impl == IMPL_FFMA else 1)
    return PackedMLP(buf, tuple(dims), impl, tuple(bw))


def _idx(t: Tensor | None) -> int | None:
    if t is None:
        return None
    if t.dtype != torch.int32 or not t.is_contiguous():
        raise TypeError("row indices must be contiguous int32")
    return t.data_ptr()


def fused_mlp(blocks: Sequence[Block], n_rows: int, packed: PackedMLP, *, final_act: int = ACT_NONE,
              act_eps: float = 0.0, res: Tensor | None = None, res_a: float = 0.0, res_b: float = 1.0,
              row_scale: Tensor | None = None, out_scale: Tensor | None = None, out: Tensor | None = None, out_index: Tensor | None = None,
              want_out: bool = True, aggr: Tensor | None = None, seg_id: Tensor | None = None,
              rowptr: Tensor | None = None, out_rows: int | None = None, gate: Tensor | None = None,
              save_hidden: list | None = None) -> Tensor | None:
    """``out[orow(r)] = epilogue(MLP(cat_s act_s(block_s[irow_s(r)])))`` -- see
    ``gtb_fused_mlp_f32`` in include/gtb200.h."""
    tensors = [_f32c(b.tensor) for b in blocks]
    dev = require_cuda(*tensors)
    d = MlpDesc()
    d.n_rows = n_rows
    d.n_srcs = len(blocks)
    d.n_layers = packed.n_layers
    if len(blocks) > _lib.GTB_MAX_SRCS:
        raise ValueError(f"at most {_lib.GTB_MAX_SRCS} column blocks")
    for i, (b, t) in enumerate(zip(blocks, tensors)):
        d.srcs[i] = Src(t.data_ptr(), _idx(b.index), t.size(1), t.stride(0), int(b.relu), b.flags)
    stream = tuple(t.size(1) for b, t in zip(blocks, tensors) if not b.projected)
    if packed.impl == IMPL_TCGEN05 and stream != packed.block_widths:
        raise ValueError(f"weights were packed for source blocks {packed.block_widths}, called with {stream}")
    for i, v in enumerate(packed.dims):
        d.dims[i] = v
    d.packed = packed.buf.data_ptr()
    d.impl = packed.impl
    d.final_act = final_act
    d.act_eps = act_eps
    n_out = packed.dims[-1]
    keep = [tensors]
    if res is not None:
        res = _f32c(res)
        d.res, d.res_ld, d.res_a, d.res_b = res.data_ptr(), res.stride(0), res_a, res_b
    d.res_b = res_b
    if row_scale is not None:
        d.row_scale = row_scale.data_ptr()
    if out_scale is not None:
        d.out_scale = out_scale.data_ptr()
    if want_out:
        if out is None:
            out = torch.empty((n_rows if out_rows is None else out_rows, n_out), dtype=torch.float32, device=dev)
        d.out, d.out_ld = out.data_ptr(), out.stride(0)
        d.out_index = _idx(out_index)
    if gate is not None:
        gate = _f32c(gate)
        d.gate, d.gate_ld = gate.data_ptr(), gate.stride(0)
    if aggr is not None:
        d.aggr, d.aggr_ld = aggr.data_ptr(), aggr.stride(0)
        d.seg_id, d.rowptr = _idx(seg_id), _idx(rowptr)
    if save_hidden is not None and n_rows > 0 and lib().gtb_fused_mlp_saves_hidden(C.byref(d)):
        # the kernel that takes this launch can hand out the two hidden activations (the backward pass then
        # does not recompute them): appended to ``save_hidden``
        h0 = torch.empty((n_rows, 64), dtype=torch.float32, device=dev)
        h1 = torch.empty((n_rows, 64), dtype=torch.float32, device=dev)
        d.hidden0, d.hidden1, d.hidden_ld = h0.data_ptr(), h1.data_ptr(), 64
        save_hidden.extend([h0, h1])
    if n_rows > 0:  # an empty edge set (E = 0) launches nothing; `out` is empty, `aggr` stays zero
        with on_device(dev):
            check(lib().gtb_fused_mlp_f32(C.byref(d), stream_ptr(dev)))
        _count(1)
    del keep
    return out if want_out else None


def in_node_fused(x: Tensor, relu_x: bool, *, aggr: Tensor | None = None, zero_aggr: bool = True,
                  packed_obj: PackedMLP | None = None, res: Tensor | None = None, res_a: float = 0.0, res_b: float = 1.0,
                  proj: tuple[PackedMLP, PackedMLP] | None = None, proj_relu: bool = False,
                  out_pa: Tensor | None = None, out_pb: Tensor | None = None):
    """Node side of a 64-wide Interaction-Network layer in one launch (``gtb_in_node_fused_f32`` in
    include/gtb200.h): returns ``(x_out, p_a, p_b)`` -- the object model's output with the residual fused in
    (None without ``packed_obj``: projection only) and the two per-node products the next consumer gathers
    (None without ``proj``).  ``aggr`` is handed back zeroed.  ``out_pa`` / ``out_pb``: tables of at least as many
    rows to write the products into (the extended table of a halo exchange: the owned rows in place)."""
    x = _f32c(x)
    dev = require_cuda(x, aggr, res)
    n = x.size(0)
    if x.size(1) != 64 or (aggr is not None and tuple(aggr.shape) != (n, 64)):
        raise ValueError("in_node_fused takes 64-column tables")
    for pk, dims in ((packed_obj, (128, 64, 64, 64)), (proj[0] if proj else None, (64, 64)), (proj[1] if proj else None, (64, 64))):
        if pk is not None and (pk.impl != IMPL_TCGEN05 or pk.dims != dims):
            raise ValueError(f"in_node_fused: weights must be tcgen05 packs of dims {dims}, got {pk.dims} (impl {pk.impl})")
    if packed_obj is not None and packed_obj.block_widths != (64, 64):
        raise ValueError("in_node_fused: the object model must be packed for the blocks (64, 64)")
    x_out = torch.empty((n, 64), dtype=torch.float32, device=dev) if packed_obj is not None else None
    p_a = p_b = None
    if proj:
        for o in (out_pa, out_pb):
            if o is not None and (o.dtype != torch.float32 or o.dim() != 2 or o.size(0) < n or o.size(1) != 64
                                  or not o.is_contiguous() or o.device != dev):
                raise ValueError("in_node_fused: out_pa / out_pb must be contiguous fp32 [>= n, 64] tables on the same device")
        p_a = out_pa if out_pa is not None else torch.empty((n, 64), dtype=torch.float32, device=dev)
        p_b = out_pb if out_pb is not None else torch.empty((n, 64), dtype=torch.float32, device=dev)
    if res is not None:
        res = _f32c(res)
    if n:
        with on_device(dev):
            check(lib().gtb_in_node_fused_f32(
                x.data_ptr(), x.stride(0), int(relu_x), aggr.data_ptr() if aggr is not None else None,
                aggr.stride(0) if aggr is not None else 0, int(zero_aggr), n,
                packed_obj.buf.data_ptr() if packed_obj is not None else None, res_a, res_b,
                res.data_ptr() if res is not None else None, res.stride(0) if res is not None else 0,
                x_out.data_ptr() if x_out is not None else None, 64,
                proj[0].buf.data_ptr() if proj else None, proj[1].buf.data_ptr() if proj else None, int(proj_relu),
                p_a.data_ptr() if proj else None, 64, p_b.data_ptr() if proj else None, 64, stream_ptr(dev)))
        _count(1)
    return x_out, p_a, p_b


def edge_encoder(x: Tensor, index: Tensor | None, n_rows: int, w0: Tensor, b0: Tensor | None, packed_w1: PackedMLP,
                 final_relu: bool) -> Tensor:
    """Two-Linear encoder 4 -> 64 -> 64 over gathered rows in one launch (``gtb_edge_encoder_f32`` in
    include/gtb200.h): ``act(W1 relu(W0 x[index] + b0) + b1)``."""
    x = _f32c(x)
    dev = require_cuda(x, index, w0, b0)
    if x.size(1) != 4 or tuple(w0.shape) != (64, 4) or packed_w1.impl != IMPL_TCGEN05 or packed_w1.dims != (64, 64):
        raise ValueError("edge_encoder takes 4 feature columns, a [64, 4] first Linear and a tcgen05 pack of a 64 -> 64 Linear")
    if x.stride(0) % 4:
        x = x.contiguous()
    w0 = w0.detach().to(torch.float32).contiguous()
    b0 = None if b0 is None else b0.detach().to(torch.float32).contiguous()
    out = torch.empty((n_rows, 64), dtype=torch.float32, device=dev)
    if n_rows:
        with on_device(dev):
            check(lib().gtb_edge_encoder_f32(x.data_ptr(), x.stride(0), _idx(index), n_rows, x.size(0), w0.data_ptr(),
                                             b0.data_ptr() if b0 is not None else None, packed_w1.buf.data_ptr(),
                                             int(final_relu), out.data_ptr(), out.stride(0), stream_ptr(dev)))
        _count(1)
    return out


def pack_in_edge_bf16(weights: Sequence[Tensor], biases: Sequence[Tensor | None]) -> Tensor:
    """Packs the relational model of a 128 / 128 / 128 Interaction-Network layer for
    ``in_edge_bf16``: three ``[128, 128]`` weights (the first one = the EDGE columns of the first
    Linear) and their biases, rounded to bf16 as ``torch.autocast`` casts them."""
    dev = require_cuda(*weights)
    if len(weights) != 3 or any(tuple(w.shape) != (128, 128) for w in weights):
        raise ValueError("in_edge_bf16 takes three [128, 128] weights")
    ws = [w.detach().to(torch.float32).contiguous() for w in weights]
    bs = [None if b is None else b.detach().to(torch.float32).contiguous() for b in biases]
    buf = torch.empty(lib().gtb_in_edge_bf16_packed_bytes(), dtype=torch.uint8, device=dev)
    wp = (C.c_void_p * 3)(*[w.data_ptr() for w in ws])
    bp = (C.c_void_p * 3)(*[(b.data_ptr() if b is not None else None) for b in bs])
    with on_device(dev):
        check(lib().gtb_in_edge_bf16_pack(wp, bp, buf.data_ptr(), stream_ptr(dev)))
    _count(1)
    return buf


def in_edge_bf16(e_in: Tensor, p_i: Tensor, p_j: Tensor, src_sorted: Tensor, dst_sorted: Tensor, packed: Tensor,
                 n_nodes: int, *, e_index: Tensor | None = None, out_index: Tensor | None = None,
                 relu_e: bool = False) -> tuple[Tensor, Tensor]:
    """bf16 Interaction-Network edge kernel (``gtb_in_edge_forward_bf16`` in include/gtb200.h):
    returns ``(e_out bf16 [E, 128], aggr fp32 [n_nodes, 128])``."""
    dev = require_cuda(e_in, p_i, p_j, src_sorted, dst_sorted, packed)
    for t in (e_in, p_i, p_j):
        if t.dtype != torch.bfloat16 or t.dim() != 2 or t.size(1) != 128 or t.stride(1) != 1:
            raise TypeError("in_edge_bf16 takes bf16 [*, 128] tables")
    n_edges = src_sorted.numel()
    e_out = torch.empty((n_edges, 128), dtype=torch.bfloat16, device=dev)
    aggr = torch.zeros((n_nodes, 128), dtype=torch.float32, device=dev)
    if n_edges:
        with on_device(dev):
            check(lib().gtb_in_edge_forward_bf16(
                e_in.data_ptr(), e_in.stride(0), _idx(e_index), int(relu_e), p_i.data_ptr(), p_i.stride(0),
                p_j.data_ptr(), p_j.stride(0), n_edges, _idx(src_sorted), _idx(dst_sorted), packed.data_ptr(),
                e_out.data_ptr(), e_out.stride(0), _idx(out_index), aggr.data_ptr(), aggr.stride(0), stream_ptr(dev)))
        _count(1)
    return e_out, aggr


def rows_atb(a: Tensor, b: Tensor, out: Tensor, *, a_index: Tensor | None = None, a_relu: bool = False,
             colsum: Tensor | None = None) -> None:
    """``out[Ka, Nb] += act(a[a_index])^T @ b`` (and ``colsum += b.sum(0)``): weight / bias gradient
    of one Linear.  Blocks wider than 64 columns are split here."""
    a, b = _f32c(a), _f32c(b)
    dev = require_cuda(a, b, out)
    n = b.size(0)
    for k0 in range(0, a.size(1), 64):
        ka = min(64, a.size(1) - k0)
        for n0 in range(0, b.size(1), 64):
            nb = min(64, b.size(1) - n0)
            cs = None
            if colsum is not None and k0 == 0:
                cs = colsum.data_ptr() + 4 * n0
            check(lib().gtb_rows_atb_f32(a.data_ptr() + 4 * k0, a.stride(0), _idx(a_index), int(a_relu), ka,
                                         b.data_ptr() + 4 * n0, b.stride(0), nb, n,
                                         out.data_ptr() + 4 * (k0 * out.stride(0) + n0), out.stride(0), cs, stream_ptr(dev)))
            _count(1)


def rows_scatter_add(src: Tensor, index: Tensor, dst: Tensor) -> None:
    """``dst[index[r]] += src[r]`` (int32 index): gradient of a row gather."""
    src = _f32c(src)
    dev = require_cuda(src, index, dst)
    if src.size(0):
        check(lib().gtb_rows_scatter_add_f32(src.data_ptr(), src.stride(0), _idx(index), src.size(0), src.size(1),
                                             dst.data_ptr(), dst.stride(0), stream_ptr(dev)))
        _count(1)


def rows_gather(table: Tensor, index: Tensor, out: Tensor | None = None) -> Tensor:
    """``out[i] = table[index[i]]`` (int32 index): packs the halo rows a peer rank needs."""
    table = _f32c(table)
    dev = require_cuda(table, index)
    n, w = index.numel(), table.size(1)
    if out is None:
        out = torch.empty((n, w), dtype=torch.float32, device=dev)
    if n:
        check(lib().gtb_rows_gather_f32(table.data_ptr(), table.stride(0), _idx(index), n, w, out.data_ptr(),
                                        out.stride(0), stream_ptr(dev)))
        _count(1)
    return out


def rows_gather_add(table: Tensor, index: Tensor, dst: Tensor) -> Tensor:
    """``dst[i] += table[index[i]]`` in place (int32 index); returns ``dst``."""
    table = _f32c(table)
    dev = require_cuda(table, index, dst)
    if dst.dtype != torch.float32 or dst.dim() != 2 or dst.stride(1) != 1 or dst.size(1) != table.size(1):
        raise TypeError("rows_gather_add: dst must be a row-major fp32 table of the source's width")
    n = index.numel()
    if n:
        check(lib().gtb_rows_gather_add_f32(table.data_ptr(), table.stride(0), _idx(index), n, table.size(1), dst.data_ptr(),
                                            dst.stride(0), stream_ptr(dev)))
        _count(1)
    return dst


def rows_inv_l2norm(blocks: Sequence[Block], n_rows: int, eps: float = 1e-12) -> Tensor:
    tensors = [_f32c(b.tensor) for b in blocks]
    dev = require_cuda(*tensors)
    arr = (Src * len(blocks))()
    for i, (b, t) in enumerate(zip(blocks, tensors)):
        arr[i] = Src(t.data_ptr(), _idx(b.index), t.size(1), t.stride(0), int(b.relu), 0)
    out = torch.empty(n_rows, dtype=torch.float32, device=dev)
    check(lib().gtb_rows_inv_l2norm_f32(arr, len(blocks), n_rows, eps, out.data_ptr(), stream_ptr(dev)))
    _count(1)
    return out
