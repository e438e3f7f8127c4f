"""linalg: general_mat_mul / general_mat_vec_mul / dot (reference src/linalg/impl_linalg.rs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import check, lib
from .array import Array


def general_mat_mul(alpha, a: Array, b: Array, beta, c: Array) -> None:
    """C <- alpha*A*B + beta*C (impl_linalg.rs:633). Shape mismatch raises ShapeError where ndarray panics."""
    al = np.asarray(alpha, dtype=c.dtype)
    be = np.asarray(beta, dtype=c.dtype)
    check(lib().ndb_gemm(c.ctx.handle, C.c_void_p(al.ctypes.data), C.byref(a.v), C.byref(b.v),
                         C.c_void_p(be.ctypes.data), C.byref(c.v)))


def general_mat_vec_mul(alpha, a: Array, x: Array, beta, y: Array) -> None:
    """y <- alpha*A*x + beta*y (impl_linalg.rs:657)."""
    al = np.asarray(alpha, dtype=y.dtype)
    be = np.asarray(beta, dtype=y.dtype)
    check(lib().ndb_gemv(y.ctx.handle, C.c_void_p(al.ctypes.data), C.byref(a.v), C.byref(x.v),
                         C.c_void_p(be.ctypes.data), C.byref(y.v)))


def dot(a: Array, b: Array):
    """Dot::dot (impl_linalg.rs:160-386)."""
    if a.ndim == 1 and b.ndim == 1:
        if a.shape != b.shape:
            raise _ffi.ShapeError(_ffi.ERR_SHAPE, "ndarray: dot product of vectors of unequal length")
        out = np.zeros((), dtype=a.dtype)
        check(lib().ndb_dot(a.ctx.handle, C.byref(a.v), C.byref(b.v), C.c_void_p(out.ctypes.data)))
        return out[()]
    if a.ndim == 2 and b.ndim == 2:
        m, k = a.shape
        k2, n = b.shape
        if k != k2:
            raise _ffi.ShapeError(_ffi.ERR_SHAPE, f"ndarray: inputs {m} × {k} and {k2} × {n} are not compatible for "
                                  "matrix multiplication")
        # output is F-order only when both inputs have unit row stride (impl_linalg.rs:319-327)
        order = "F" if (a.strides[0] == 1 and b.strides[0] == 1) else "C"
        c = Array.uninit((m, n), a.dtype, order, a.ctx)
        general_mat_mul(1, a, b, 0, c)
        return c
    if a.ndim == 2 and b.ndim == 1:
        m, k = a.shape
        if k != b.shape[0]:
            raise _ffi.ShapeError(_ffi.ERR_SHAPE, f"ndarray: inputs {m} × {k} and {b.shape[0]} × 1 are not compatible "
                                  "for matrix multiplication")
        y = Array.uninit((m,), a.dtype, "C", a.ctx)
        general_mat_vec_mul(1, a, b, 0, y)
        return y
    if a.ndim == 1 and b.ndim == 2:
        # vector · matrix = (matrixᵀ · vector) (impl_linalg.rs:245-262)
        return dot(b.t(), a)
    raise _ffi.NdbError(_ffi.ERR_UNSUPPORTED, "dot: operands must be 1-d or 2-d")
