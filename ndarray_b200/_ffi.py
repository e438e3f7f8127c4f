"""ctypes binding of libndb200.so (include/ndb200.h).

This is what a host in any language does: dlopen the C-ABI library and call it with plain pointers.
There is no fallback: if the library cannot be loaded, importing the compute API raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

MAX_NDIM = 8
MAX_PARTS = 9

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("NDB200_LIB", HERE / "libndb200.so"))


class View(C.Structure):
    """ndb_view: (ptr, ndim, dtype, dim[8], stride[8]) — what NdProducer exposes."""
    _fields_ = [("ptr", C.c_void_p), ("ndim", C.c_int32), ("dtype", C.c_int32),
                ("dim", C.c_int64 * MAX_NDIM), ("stride", C.c_int64 * MAX_NDIM)]

    def copy(self) -> "View":
        v = View()
        C.memmove(C.byref(v), C.byref(self), C.sizeof(View))
        return v


class ZipPlan(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("ndim", C.c_int32), ("dim", C.c_int64 * MAX_NDIM),
                ("stride", (C.c_int64 * MAX_NDIM) * MAX_PARTS), ("offset", C.c_int64 * MAX_PARTS),
                ("tile_axis", C.c_int32), ("staged_mask", C.c_uint32)]


# status codes (include/ndb200.h)
OK, ERR_SHAPE, ERR_LAYOUT, ERR_RANGE, ERR_OOB, ERR_UNSUPPORTED, ERR_OVERFLOW, ERR_INVALID, ERR_EMPTY, ERR_DIVZERO, \
    ERR_ALIASING = range(11)
ERR_CUDA = 100

KERNEL_NAMES = {0: "noop", 1: "flat_vec", 2: "flat", 3: "rows_vec", 4: "rows", 5: "tile"}


class NdbError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[ndb status {status}] {message}")
        self.status = status
        self.message = message


class ShapeError(NdbError):
    """ndarray panics / returns ShapeError(IncompatibleShape)."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ndarray_b200.build` (nvcc, sm_100a). "
            "ndarray_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    PV = C.POINTER(View)
    sigs = {
        "ndb_version": (C.c_char_p, []),
        "ndb_last_error": (C.c_char_p, []),
        "ndb_launch_count": (u64, []),
        "ndb_device_count": (i32, [C.POINTER(i32)]),
        "ndb_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "ndb_ctx_destroy": (i32, [vp]),
        "ndb_ctx_set_stream": (i32, [vp, vp]),
        "ndb_ctx_stream": (vp, [vp]),
        "ndb_sync": (i32, [vp]),
        "ndb_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "ndb_free": (i32, [vp, vp]),
        "ndb_memset": (i32, [vp, vp, i32, C.c_size_t]),
        "ndb_h2d": (i32, [vp, vp, vp, C.c_size_t]),
        "ndb_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "ndb_d2d": (i32, [vp, vp, vp, C.c_size_t]),
        "ndb_host_alloc": (i32, [C.c_size_t, C.POINTER(vp)]),
        "ndb_host_free": (i32, [vp]),
        "ndb_graph_begin": (i32, [vp]),
        "ndb_graph_end": (i32, [vp, C.POINTER(vp)]),
        "ndb_graph_launch": (i32, [vp, vp]),
        "ndb_graph_destroy": (i32, [vp]),
        "ndb_event_create": (i32, [C.POINTER(vp)]),
        "ndb_event_destroy": (i32, [vp]),
        "ndb_event_record": (i32, [vp, vp]),
        "ndb_event_elapsed_ms": (i32, [vp, vp, C.POINTER(C.c_float)]),
        "ndb_view_slice_axis": (i32, [PV, i32, i64, i64, i32, i64]),
        "ndb_view_index_axis": (i32, [PV, i32, i64]),
        "ndb_view_broadcast": (i32, [PV, i32, C.POINTER(i64), PV]),
        "ndb_view_permute": (i32, [PV, C.POINTER(i32), PV]),
        "ndb_view_invert_axis": (i32, [PV, i32]),
        "ndb_view_split_at": (i32, [PV, i32, i64, PV, PV]),
        "ndb_view_merge_axes": (i32, [PV, i32, i32]),
        "ndb_view_collapse_axis": (i32, [PV, i32, i64]),
        "ndb_view_layout": (C.c_uint32, [PV]),
        "ndb_view_is_contiguous": (i32, [PV]),
        "ndb_view_min_stride_axis": (i32, [PV]),
        "ndb_zip": (i32, [vp, i32, vp, i32, PV]),
        "ndb_cast": (i32, [vp, PV, PV]),
        "ndb_zip_plan": (i32, [i32, i32, PV, C.POINTER(ZipPlan)]),
        "ndb_zip_prefer_f": (i32, [i32, PV]),
        "ndb_zip_host": (i32, [vp, i32, vp, i32, PV]),
        "ndb_zip_host_fold": (i32, [vp, i32, vp, i32, PV, i32, vp]),
        "ndb_fill_counter_rng": (i32, [vp, PV, u64, u64]),
        "ndb_zip_all": (i32, [vp, i32, vp, i32, PV, C.POINTER(i32)]),
        "ndb_zip_any": (i32, [vp, i32, vp, i32, PV, C.POINTER(i32)]),
        "ndb_zip_fold_sum": (i32, [vp, i32, vp, i32, PV, vp]),
        "ndb_zip_fold": (i32, [vp, i32, vp, i32, PV, i32, vp, vp]),
        "ndb_zip_fold_write": (i32, [vp, i32, vp, i32, PV, i32, vp, i32]),
        "ndb_zip_fold_while": (i32, [vp, i32, vp, i32, PV, i32, vp, i32, vp, vp, C.POINTER(i32)]),
        "ndb_var": (i32, [vp, PV, dbl, i32, vp]),
        "ndb_reduce_all": (i32, [vp, i32, PV, vp]),
        "ndb_reduce_all_dev": (i32, [vp, i32, PV, vp]),
        "ndb_reduce_axis": (i32, [vp, i32, i32, PV, PV, dbl]),
        "ndb_welford_axis": (i32, [vp, i32, PV, PV, PV]),
        "ndb_welford_merge": (i32, [vp, i32, C.POINTER(i64), PV, PV, dbl, i32, PV]),
        "ndb_peer_window_create": (i32, [vp, i32, i32, C.c_size_t, vp]),
        "ndb_peer_window_attach": (i32, [vp, vp]),
        "ndb_peer_window_destroy": (i32, [vp]),
        "ndb_peer_allreduce": (i32, [vp, i32, i32, vp, vp, C.c_int64]),
        "ndb_peer_allgather": (i32, [vp, i32, vp, vp, C.c_int64]),
        "ndb_reduce_all_peer": (i32, [vp, i32, PV, vp]),
        "ndb_reduce_axis_peer": (i32, [vp, i32, i32, PV, PV, C.c_double, C.c_int64]),
        "ndb_reduce_axis_peer_ok": (i32, [vp, i32, i32, PV, PV, C.c_double, C.c_int64]),
        "ndb_comm_unique_id": (i32, [vp]),
        "ndb_comm_init": (i32, [vp, i32, i32, vp]),
        "ndb_comm_destroy": (i32, [vp]),
        "ndb_broadcast": (i32, [vp, PV, i32]),
        "ndb_gemm_rowsharded": (i32, [vp, vp, PV, PV, vp, PV, i32]),
        "ndb_reduce_axis_sharded": (i32, [vp, i32, i32, PV, PV, dbl, C.c_int64]),
        "ndb_reduce_all_sharded": (i32, [vp, i32, PV, vp]),
        "ndb_accumulate_axis": (i32, [vp, i32, i32, PV]),
        "ndb_dot": (i32, [vp, PV, PV, vp]),
        "ndb_gemm": (i32, [vp, vp, PV, PV, vp, PV]),
        "ndb_gemv": (i32, [vp, vp, PV, PV, vp, PV]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status: int) -> None:
    if status == OK:
        return
    msg = lib().ndb_last_error().decode("utf-8", "replace")
    if status == ERR_SHAPE:
        raise ShapeError(status, msg)
    raise NdbError(status, msg)
