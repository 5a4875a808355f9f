"""ctypes binding of ``libgtb200.so`` (``include/gtb200.h``).

The library is the product: there is no PyTorch / CPU fallback.  If it is missing
or the device is not a compute-capability 10.x GPU every op raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libgtb200.so"

GTB_MAX_SRCS = 16
GTB_MAX_LAYERS = 3
GTB_MAX_WIDTH = 128
ACT_NONE, ACT_RELU, ACT_SIGMOID_AFFINE = 0, 1, 2
IMPL_AUTO, IMPL_FFMA, IMPL_TCGEN05 = 0, 1, 2
SRC_PROJECTED, SRC_SORTED = 1, 2

_ERR_NAMES = {-1: "BAD_ARG", -2: "UNSUPPORTED_DIM", -3: "WORKSPACE", -4: "CUDA", -5: "ARCH"}


class GtbError(RuntimeError):
    """Error reported by libgtb200.  CUDA allocation failures keep the words
    "out of memory" in the message so that the reference's
    ``tolerate_some_oom_errors`` (utils/oom.py:12-18) still recognises them."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"gtb200 {_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("index", C.c_void_p), ("width", C.c_int32), ("ld", C.c_int32),
                ("relu", C.c_int32), ("flags", C.c_int32)]


class MlpDesc(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("n_srcs", C.c_int32), ("n_layers", C.c_int32),
        ("srcs", Src * GTB_MAX_SRCS), ("dims", C.c_int32 * (GTB_MAX_LAYERS + 1)),
        ("packed", C.c_void_p), ("impl", C.c_int32), ("final_act", C.c_int32), ("act_eps", C.c_float),
        ("res_a", C.c_float), ("res_b", C.c_float), ("res_ld", C.c_int32), ("res", C.c_void_p),
        ("row_scale", C.c_void_p), ("out_scale", C.c_void_p),
        ("out", C.c_void_p), ("out_index", C.c_void_p), ("out_ld", C.c_int32),
        ("aggr_ld", C.c_int32), ("aggr", C.c_void_p), ("seg_id", C.c_void_p), ("rowptr", C.c_void_p),
        ("gate", C.c_void_p), ("gate_ld", C.c_int32), ("hidden_ld", C.c_int32),
        ("hidden0", C.c_void_p), ("hidden1", C.c_void_p),
    ]


_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/gtb200.h declares
SIGNATURES = {
    "gtb_version": (C.c_int, []),
    "gtb_last_error": (C.c_char_p, []),
    "gtb_arch_ok": (C.c_int, [C.c_int]),
    "gtb_plan_workspace_bytes": (_sz, [_i64, _i64]),
    "gtb_plan_build": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gtb_plan_filter_workspace_bytes": (_sz, [_i64, _i64]),
    "gtb_plan_filter": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gtb_plan_prune_workspace_bytes": (_sz, [_i64]),
    "gtb_plan_prune_orphans": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gtb_mlp_packed_bytes": (_sz, [C.c_int, C.POINTER(_i32), C.c_int, C.POINTER(_i32), C.c_int]),
    "gtb_mlp_pack": (C.c_int, [C.c_int, C.POINTER(_i32), C.c_int, C.POINTER(_i32), C.POINTER(_vp), C.POINTER(_vp),
                               C.c_int, _vp, _vp]),
    "gtb_mlp_tc_slots": (C.c_int, [C.c_int, C.POINTER(_i32), C.c_int, C.POINTER(_i32)]),
    "gtb_fused_mlp_f32": (C.c_int, [C.POINTER(MlpDesc), _vp]),
    "gtb_fused_mlp_saves_hidden": (C.c_int, [C.POINTER(MlpDesc)]),
    "gtb_debug_tc_timeout": (C.c_int, [C.POINTER(C.c_int)]),
    "gtb_debug_tc_profile": (C.c_int, [C.c_int, C.POINTER(C.c_longlong)]),
    "gtb_in_edge_forward_f32": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp,
                                          _i32, _i32, _i32, _i32, _vp, C.c_int, _vp, _i32, _vp, _vp]),
    "gtb_in_node_forward_f32": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _vp, C.c_int,
                                          _f32, _f32, _vp, _i32, _vp, _i32, _vp]),
    "gtb_in_node_fused_f32": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _vp, _f32, _f32, _vp, _i32, _vp, _i32, _vp, _vp,
                                        _i32, _vp, _i32, _vp, _i32, _vp]),
    "gtb_edge_encoder_f32": (C.c_int, [_vp, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _i32, _vp]),
    "gtb_in_edge_bf16_packed_bytes": (_sz, []),
    "gtb_in_edge_bf16_pack": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), _vp, _vp]),
    "gtb_in_edge_forward_bf16": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _vp,
                                           _vp, _i32, _vp]),
    "gtb_ec_loss_f32": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, _f32, C.c_int, _f32, _f32, _f32, _vp, _vp]),
    "gtb_ec_loss_grad_f32": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, _f32, C.c_int, _f32, _f32, _f32, _vp, _vp, _vp]),
    "gtb_oc_workspace_bytes": (_sz, [_i64]),
    "gtb_oc_prepare": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gtb_oc_alphas": (C.c_int, [_vp, _vp, _i64, _f32, _i32, _vp, _vp, _vp]),
    "gtb_oc_potentials": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _i32, _f32, _i64, _vp, _vp]),
    "gtb_radius_pair_sum_f32": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _i32, _i32, _vp, _vp]),
    "gtb_radius_graph_count_f32": (C.c_int, [_vp, _i32, _i64, _vp, _f32, _i32, _i32, _vp, _vp]),
    "gtb_radius_graph_fill_f32": (C.c_int, [_vp, _i32, _i64, _vp, _f32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "gtb_radius_graph_grid_workspace_bytes": (_sz, [_i64]),
    "gtb_radius_pair_sum_grid_f32": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "gtb_radius_pair_sum_grad_grid_f32": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _i32, _i32, _vp, _vp, _vp,
                                                     _vp, _sz, _vp]),
    "gtb_radius_graph_grid_count_f32": (C.c_int, [_vp, _i32, _i64, _vp, _f32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "gtb_radius_graph_grid_fill_f32": (C.c_int, [_vp, _i32, _i64, _vp, _f32, _i32, _i32, _vp, _vp, _i64, _vp, _sz, _vp]),
    "gtb_radius_pair_sum_grad_f32": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "gtb_edge_dist_pow_grad_f32": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _f32, _vp, _vp, _vp]),
    "gtb_edge_dist_pow_sum_f32": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _f32, _vp, _vp]),
    "gtb_oc_potentials_grad": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i64, _vp, _i32, _f32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "gtb_dbscan_f32": (C.c_int, [_vp, _i32, _i64, C.c_double, _i32, _vp, _vp, _vp, _vp]),
    "gtb_dbscan_grid_workspace_bytes": (_sz, [_i64]),
    "gtb_dbscan_grid_f32": (C.c_int, [_vp, _i32, _i64, C.c_double, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gtb_rows_inv_l2norm_f32": (C.c_int, [C.POINTER(Src), _i32, _i64, _f32, _vp, _vp]),
    "gtb_rows_atb_f32": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _vp]),
    "gtb_rows_scatter_add_f32": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _i32, _vp]),
    "gtb_rows_gather_f32": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _i32, _vp]),
    "gtb_rows_scatter_f32": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _i32, _vp]),
    "gtb_rows_gather_add_f32": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _i32, _vp]),
}


@functools.lru_cache(maxsize=1)
def lib() -> C.CDLL:
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m gnn_tracking_b200.csrc.build` "
            "(gnn_tracking_b200 has no fallback implementation)")
    handle = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    return handle


def check(rc: int) -> None:
    if rc != 0:
        raise GtbError(rc, lib().gtb_last_error().decode(errors="replace"))


@functools.lru_cache(maxsize=None)
def require_device(index: int) -> None:
    check(lib().gtb_arch_ok(index))
