"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE — see ndarray_oracle.cpp's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Views are numpy arrays: a numpy array's (data pointer, shape, strides/itemsize) IS ndarray's
(ptr, dim, strides) triple, negative strides included, so the oracle runs on exactly the view the
device path is given.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libndarray_oracle.so"
MAX_NDIM = 8

DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}


class OrcView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ndim", C.c_int32), ("dtype", C.c_int32),
                ("dim", C.c_int64 * MAX_NDIM), ("stride", C.c_int64 * MAX_NDIM)]


def build(force: bool = False) -> Path:
    src = HERE / "ndarray_oracle.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_layout.restype = C.c_uint32
    return _lib


def view(a: np.ndarray) -> OrcView:
    """Describe a numpy array (any strides) as an ndarray-style view."""
    if a.dtype not in DTYPES:
        raise TypeError(a.dtype)
    if a.ndim > MAX_NDIM:
        raise ValueError("rank")
    v = OrcView()
    v.ptr = a.ctypes.data if a.size else (a.ctypes.data or 0)
    v.ndim = a.ndim
    v.dtype = DTYPES[a.dtype]
    for i in range(a.ndim):
        v.dim[i] = a.shape[i]
        assert a.strides[i] % a.itemsize == 0
        v.stride[i] = a.strides[i] // a.itemsize
    return v


def raw_view(ptr: int, dtype, dim, stride) -> OrcView:
    v = OrcView()
    v.ptr = ptr
    v.ndim = len(dim)
    v.dtype = DTYPES[np.dtype(dtype)]
    for i, (d, s) in enumerate(zip(dim, stride)):
        v.dim[i] = d
        v.stride[i] = s
    return v


def _scal(dtype, vals):
    if not vals:
        return None
    arr = np.asarray(vals, dtype=dtype)
    return arr


def zip_(op: int, out: np.ndarray, inputs, scalars=(), threads: int = 1) -> int:
    """out <- op(inputs...) with ndarray's Zip traversal. Returns the oracle status (0 = ok)."""
    parts = (OrcView * (1 + len(inputs)))(view(out), *[view(x) for x in inputs])
    sc = _scal(out.dtype, scalars)
    return lib().orc_zip(op, sc.ctypes.data if sc is not None else None, len(parts), parts, threads)


def reduce_all(redop: int, a: np.ndarray):
    out = np.zeros((), dtype=a.dtype)
    v = view(a)
    st = lib().orc_reduce_all(redop, C.byref(v), C.c_void_p(out.ctypes.data))
    if st:
        raise RuntimeError(f"oracle status {st}")
    return out[()]


def reduce_axis(redop: int, a: np.ndarray, axis: int, ddof: float = 0.0):
    """Returns (status, out). status 8 == ndarray's None, 7 == ndarray panics."""
    shape = a.shape[:axis] + a.shape[axis + 1:]
    out = np.zeros(shape, dtype=a.dtype)
    va, vo = view(a), view(out)
    st = lib().orc_reduce_axis(redop, axis, C.byref(va), C.byref(vo), C.c_double(ddof))
    return st, out


def var(a: np.ndarray, ddof: float):
    out = C.c_double()
    v = view(a)
    st = lib().orc_var(C.byref(v), C.c_double(ddof), C.byref(out))
    return st, out.value


def dot(x: np.ndarray, y: np.ndarray):
    out = np.zeros((), dtype=x.dtype)
    vx, vy = view(x), view(y)
    st = lib().orc_dot(C.byref(vx), C.byref(vy), C.c_void_p(out.ctypes.data))
    if st:
        raise RuntimeError(f"oracle status {st}")
    return out[()]


def sum_exact(a: np.ndarray) -> float:
    out = C.c_double()
    v = view(a)
    st = lib().orc_sum_exact(C.byref(v), C.byref(out))
    if st:
        raise RuntimeError(f"oracle status {st}")
    return out.value


def gemm(alpha, a: np.ndarray, b: np.ndarray, beta, c: np.ndarray, mode: int = 0, threads: int = 1) -> int:
    """c <- alpha*a@b + beta*c in place. mode 0 = naive fold (tests/oper.rs:291), 1 = Kahan (accuracy.rs:25)."""
    al = np.asarray(alpha, dtype=a.dtype)
    be = np.asarray(beta, dtype=a.dtype)
    va, vb, vc = view(a), view(b), view(c)
    return lib().orc_gemm(mode, C.c_void_p(al.ctypes.data), C.byref(va), C.byref(vb), C.c_void_p(be.ctypes.data),
                          C.byref(vc), threads)


def slice_axis(v: OrcView, axis: int, start: int, end, step: int) -> int:
    return lib().orc_slice_axis(C.byref(v), axis, start, 0 if end is None else end, 0 if end is None else 1, step)


def layout(a) -> int:
    v = a if isinstance(a, OrcView) else view(a)
    return lib().orc_layout(C.byref(v))


def is_contiguous(a) -> bool:
    v = a if isinstance(a, OrcView) else view(a)
    return bool(lib().orc_is_contiguous(C.byref(v)))


def min_stride_axis(a) -> int:
    v = a if isinstance(a, OrcView) else view(a)
    return lib().orc_min_stride_axis(C.byref(v))


def max_stride_axis(a) -> int:
    v = a if isinstance(a, OrcView) else view(a)
    return lib().orc_max_stride_axis(C.byref(v))


def zip_prefer_f(inputs) -> bool:
    parts = (OrcView * len(inputs))(*[x if isinstance(x, OrcView) else view(x) for x in inputs])
    return bool(lib().orc_zip_prefer_f(len(parts), parts))


def max_threads() -> int:
    """Host threads the CPU arm may use: the cores this process is allowed on.  Deliberately NOT omp_get_max_threads():
    torchrun exports OMP_NUM_THREADS=1 to every rank; the oracle's parallel loops carry explicit num_threads clauses."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, lib().orc_hw_threads()))


def counter_fill(out: np.ndarray, seed: int, first: int = 0, threads: int = 1) -> np.ndarray:
    """out.flat[j] = value(seed, first + j): the harness's index-addressable generator (csrc/rng.cu)."""
    assert out.flags.c_contiguous
    st = lib().orc_counter_fill(C.c_void_p(out.ctypes.data), DTYPES[out.dtype], C.c_int64(out.size), C.c_uint64(seed),
                                C.c_uint64(first), threads)
    if st:
        raise RuntimeError(f"oracle status {st}")
    return out


def par_row_sum(a: np.ndarray, threads: int = 1):
    """a.axis_iter(Axis(0)).into_par_iter().map(|r| r.sum()).sum()  (tests/par_rayon.rs:19)."""
    out = np.zeros((), dtype=a.dtype)
    v = view(a)
    st = lib().orc_par_row_sum(C.byref(v), threads, C.c_void_p(out.ctypes.data))
    if st:
        raise RuntimeError(f"oracle status {st}")
    return out[()]


def c5_sums(seed_a: int, seed_b: int, seed_c: int, first: int, n: int, threads: int = 1, ndarray_order: bool = False):
    """Config 5 from the counter generator, nothing materialised: (sum in f64, sum of |.| in f64, ndarray-order f32 sum|None)."""
    s, sa, f = C.c_double(), C.c_double(), C.c_float()
    st = lib().orc_c5_sums(C.c_uint64(seed_a), C.c_uint64(seed_b), C.c_uint64(seed_c), C.c_uint64(first), C.c_int64(n), threads,
                           C.byref(s), C.byref(sa), int(ndarray_order), C.byref(f))
    if st:
        raise RuntimeError(f"oracle status {st}")
    return s.value, sa.value, (np.float32(f.value) if ndarray_order else None)


def cast(out: np.ndarray, a: np.ndarray) -> int:
    vo, va = view(out), view(a)
    return lib().orc_cast(C.byref(vo), C.byref(va))
