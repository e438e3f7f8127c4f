"""Host-side mirror of ndarray's API for the hot path, over libndb200's C ABI.

The names, argument meanings and failure behaviour follow rust-ndarray 0.17.2 so that the parity
tests read like ndarray's own (`Zip.from_(a).and_(b).for_each(...)`, `a.sum_axis(Axis(0))`,
`a.slice(s_[::-2, :, :])`, `general_mat_mul(alpha, a, b, beta, c)`).  This layer owns no arithmetic
and no traversal logic: every view rewrite and every computation is one call into the C library
(view arithmetic = ndb_view_*, planning + kernels = ndb_zip / ndb_reduce_* / ndb_gemm).  A Rust
`DeviceArray<A, D>` crate would make exactly the same calls.

Where ndarray panics we raise (ShapeError / NdbError with ndarray's message); where it returns
`None` we return None.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import View, check, lib

# ---- enums shared with include/ndb200.h ---------------------------------------------------
DTYPE_CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}


class Op:
    COPY, FILL, ADD, SUB, MUL, DIV, REM, MUL_ADD, ADD_S, SUB_S, MUL_S, DIV_S, REM_S, RSUB_S, RDIV_S, NEG, AXPY, \
        ADD_N, ABS, SQRT, MIN, MAX, SCALE_ADD, TF32_HI, TF32_LO, EQ, NE, LT, LE, GT, GE, MATH, CLAMP = range(33)


MATH_FNS = ["floor", "ceil", "round", "trunc", "fract", "signum", "recip", "exp", "exp2", "exp_m1", "ln", "log2", "log10",
            "ln_1p", "cbrt", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh",
            "to_degrees", "to_radians", "pow2", "powi", "powf", "log", "abs_sub", "hypot", "is_nan", "is_infinite"]
MATH_FN = {name: i for i, name in enumerate(MATH_FNS)}


class Red:
    SUM, PROD, MIN, MAX, MEAN, VAR, STD = range(7)


class Axis(int):
    """ndarray::Axis (src/dimension/axis.rs)."""

    def index(self) -> int:
        return int(self)


class Slice:
    """ndarray::Slice {start, end: Option, step} (src/slice.rs:42).

    NOTE the semantics are ndarray's, not Python's: the range start..end is taken first, THEN a negative
    step walks it from the back (src/slice.rs:780-808): `s![1..3;-1]` of [0,1,2,3] is [2,1]."""

    def __init__(self, start: int = 0, end: Optional[int] = None, step: int = 1):
        self.start, self.end, self.step = int(start), (None if end is None else int(end)), int(step)

    def __repr__(self):
        return f"Slice({self.start}, {self.end}, {self.step})"


class _NewAxis:
    def __repr__(self):
        return "NewAxis"


NewAxis = _NewAxis()


class _SliceMaker:
    """`s_[a:b:step, i, NewAxis, ...]` — the `s![]` macro (src/slice.rs:810). Python slice syntax, ndarray meaning."""

    def __getitem__(self, item):
        if not isinstance(item, tuple):
            item = (item,)
        out = []
        for it in item:
            if isinstance(it, slice):
                out.append(Slice(0 if it.start is None else it.start, it.stop, 1 if it.step is None else it.step))
            elif isinstance(it, (Slice, _NewAxis)):
                out.append(it)
            elif isinstance(it, (int, np.integer)):
                out.append(int(it))
            else:
                raise TypeError(f"s_[]: unsupported element {it!r}")
        return tuple(out)


s_ = _SliceMaker()


# ---- context --------------------------------------------------------------------------------
class Context:
    """One device + stream + workspace (ndb_ctx)."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib().ndb_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device

    def sync(self):
        check(lib().ndb_sync(self.handle))

    def set_stream(self, cuda_stream: int | None):
        check(lib().ndb_ctx_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def stream_ptr(self) -> int:
        return int(lib().ndb_ctx_stream(self.handle) or 0)

    def close(self):
        if self.handle:
            lib().ndb_ctx_destroy(self.handle)
            self.handle = None

    # CUDA graphs: `with ctx.capture() as g: ...calls...` then g.launch() replays them with one host call
    def capture(self):
        return _Capture(self)

    # event timing on the context stream
    def event(self):
        e = C.c_void_p()
        check(lib().ndb_event_create(C.byref(e)))
        return e

    def record(self, ev):
        check(lib().ndb_event_record(self.handle, ev))

    @staticmethod
    def elapsed_ms(start, stop) -> float:
        ms = C.c_float()
        check(lib().ndb_event_elapsed_ms(start, stop, C.byref(ms)))
        return ms.value


class _Capture:
    def __init__(self, ctx: Context):
        self.ctx, self.graph = ctx, None

    def __enter__(self):
        check(lib().ndb_graph_begin(self.ctx.handle))
        return self

    def __exit__(self, et, ev, tb):
        g = C.c_void_p()
        st = lib().ndb_graph_end(self.ctx.handle, C.byref(g))
        if et is None:
            check(st)
            self.graph = g
        return False

    def launch(self):
        check(lib().ndb_graph_launch(self.ctx.handle, self.graph))

    def close(self):
        if self.graph:
            lib().ndb_graph_destroy(self.graph)
            self.graph = None


_default_ctx: dict[int, Context] = {}


def default_context() -> Context:
    dev = int(os.environ.get("NDB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if dev not in _default_ctx:
        _default_ctx[dev] = Context(dev)
    return _default_ctx[dev]


class _Buffer:
    """Owning device allocation (the OwnedRepr analogue, src/data_repr.rs:24)."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        p = C.c_void_p()
        check(lib().ndb_alloc(ctx.handle, nbytes, C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr and self.ctx.handle:
                lib().ndb_free(self.ctx.handle, C.c_void_p(self.ptr))
        except Exception:
            pass
        self.ptr = None


def _c_strides(shape):
    st, c = [0] * len(shape), 1
    nonempty = all(d != 0 for d in shape)
    for i in range(len(shape) - 1, -1, -1):
        st[i] = c if nonempty else 0
        c *= shape[i]
    return st


def _f_strides(shape):
    st, c = [0] * len(shape), 1
    nonempty = all(d != 0 for d in shape)
    for i in range(len(shape)):
        st[i] = c if nonempty else 0
        c *= shape[i]
    return st


def _as_shape(shape) -> tuple:
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(d) for d in shape)


# ---- the array / view type --------------------------------------------------------------------
class Array:
    """Device array or view: (ptr, dim, strides) over a device buffer.

    Owned arrays and views share this class (ndarray's ArrayBase<S, D>, src/lib.rs:1304); `owned` tells
    whether `self` is the sole logical owner of its elements."""

    __array_priority__ = 1000

    def __init__(self, ctx: Context, buf: _Buffer, view: View, dtype: np.dtype, owned: bool):
        self.ctx, self.buf, self.v, self.dtype, self.owned = ctx, buf, view, np.dtype(dtype), owned

    # -- construction
    @staticmethod
    def _alloc(shape, dtype, order="C", ctx: Context | None = None) -> "Array":
        ctx = ctx or default_context()
        shape = _as_shape(shape)
        dtype = np.dtype(dtype)
        if dtype not in DTYPE_CODE:
            raise TypeError(f"unsupported element type {dtype}")
        if len(shape) > _ffi.MAX_NDIM:
            raise NdbErrorUnsupported(f"rank {len(shape)} > {_ffi.MAX_NDIM}")
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        buf = _Buffer(ctx, n * dtype.itemsize)
        v = View()
        v.ptr = buf.ptr
        v.ndim = len(shape)
        v.dtype = DTYPE_CODE[dtype]
        st = _f_strides(shape) if order == "F" else _c_strides(shape)
        for i, d in enumerate(shape):
            v.dim[i] = d
            v.stride[i] = st[i]
        return Array(ctx, buf, v, dtype, True)

    @staticmethod
    def uninit(shape, dtype=np.float32, order="C", ctx=None) -> "Array":
        return Array._alloc(shape, dtype, order, ctx)

    @staticmethod
    def zeros(shape, dtype=np.float32, order="C", ctx=None) -> "Array":
        a = Array._alloc(shape, dtype, order, ctx)
        check(lib().ndb_memset(a.ctx.handle, C.c_void_p(a.buf.ptr), 0, a.buf.nbytes))
        return a

    @staticmethod
    def from_elem(shape, elem, dtype=None, order="C", ctx=None) -> "Array":
        dtype = np.dtype(dtype) if dtype is not None else np.asarray(elem).dtype
        a = Array._alloc(shape, dtype, order, ctx)
        a.fill(elem)
        return a

    @staticmethod
    def ones(shape, dtype=np.float32, order="C", ctx=None) -> "Array":
        return Array.from_elem(shape, 1, dtype, order, ctx)

    @staticmethod
    def from_numpy(a: np.ndarray, ctx=None) -> "Array":
        """Upload. C- or F-contiguous inputs keep their layout, anything else is uploaded in C order."""
        a = np.asarray(a)
        if a.ndim >= 2 and a.flags.f_contiguous and not a.flags.c_contiguous:
            out = Array._alloc(a.shape, a.dtype, "F", ctx)
            src = np.ascontiguousarray(a.T)
        else:
            out = Array._alloc(a.shape, a.dtype, "C", ctx)
            src = np.ascontiguousarray(a)
        if src.size:
            check(lib().ndb_h2d(out.ctx.handle, C.c_void_p(out.buf.ptr), C.c_void_p(src.ctypes.data), src.nbytes))
            out.ctx.sync()
        return out

    @staticmethod
    def borrow(ptr: int, shape, dtype, order="C", ctx=None, owner=None, nbytes=None) -> "Array":
        """A non-owning array over device memory someone else allocated (e.g. a torch tensor's data_ptr()).
        `owner` is kept alive for the lifetime of the array."""
        ctx = ctx or default_context()
        shape = _as_shape(shape)
        dtype = np.dtype(dtype)

        class _Borrowed:
            pass

        buf = _Borrowed()
        buf.ptr, buf.ctx, buf.owner = int(ptr), ctx, owner
        buf.nbytes = nbytes if nbytes is not None else int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        v = View()
        v.ptr = int(ptr)
        v.ndim = len(shape)
        v.dtype = DTYPE_CODE[dtype]
        st = _f_strides(shape) if order == "F" else _c_strides(shape)
        for i, d in enumerate(shape):
            v.dim[i] = d
            v.stride[i] = st[i]
        return Array(ctx, buf, v, dtype, False)

    def raw_view_with(self, offset: int, shape, strides) -> "Array":
        """ArrayView::from_shape_ptr analogue (src/impl_views/constructors.rs): a view over this array's
        buffer starting `offset` ELEMENTS after the buffer start, with explicit shape and signed element
        strides.  The caller vouches that every reachable element lies inside the buffer."""
        shape, strides = _as_shape(shape), tuple(int(x) for x in strides)
        lo = hi = int(offset)
        for d, st in zip(shape, strides):
            if d == 0:
                lo, hi = 0, -1
                break
            if st < 0:
                lo += (d - 1) * st
            else:
                hi += (d - 1) * st
        if hi >= lo and (lo < 0 or (hi + 1) * self.dtype.itemsize > max(self.buf.nbytes, 1)):
            raise _ffi.NdbError(_ffi.ERR_OOB, "raw_view_with: view reaches outside the allocation")
        v = View()
        v.ptr = self.buf.ptr + int(offset) * self.dtype.itemsize
        v.ndim = len(shape)
        v.dtype = DTYPE_CODE[self.dtype]
        for i, (d, st) in enumerate(zip(shape, strides)):
            v.dim[i] = d
            v.stride[i] = st
        return self._derive(v)

    # -- shape
    @property
    def ndim(self) -> int:
        return self.v.ndim

    @property
    def shape(self) -> tuple:
        return tuple(self.v.dim[i] for i in range(self.v.ndim))

    dim = shape

    @property
    def strides(self) -> tuple:
        """Signed strides in ELEMENTS, like ArrayBase::strides()."""
        return tuple(self.v.stride[i] for i in range(self.v.ndim))

    def len(self) -> int:
        n = 1
        for d in self.shape:
            n *= d
        return n

    def len_of(self, axis: int) -> int:
        return self.shape[axis]

    def raw_view(self) -> View:
        return self.v.copy()

    def is_standard_layout(self) -> bool:
        return self.strides == tuple(_c_strides(self.shape)) or bool(lib().ndb_view_layout(C.byref(self.v)) & 1)

    def is_contiguous(self) -> bool:
        return bool(lib().ndb_view_is_contiguous(C.byref(self.v)))

    def _derive(self, v: View) -> "Array":
        return Array(self.ctx, self.buf, v, self.dtype, False)

    def view(self) -> "Array":
        return self._derive(self.v.copy())

    # -- view transforms (all zero-copy; src/impl_methods.rs)
    def slice_axis(self, axis: int, sl: Slice) -> "Array":
        v = self.v.copy()
        check(lib().ndb_view_slice_axis(C.byref(v), int(axis), sl.start, 0 if sl.end is None else sl.end,
                                        0 if sl.end is None else 1, sl.step))
        return self._derive(v)

    def slice(self, info: Sequence) -> "Array":
        """`a.slice(s![..])` (src/impl_methods.rs:514-621): Slice keeps the axis, an int index removes it,
        NewAxis inserts a length-1 axis."""
        n_in = sum(1 for e in info if not isinstance(e, _NewAxis))
        if n_in != self.ndim:
            raise _ffi.NdbError(_ffi.ERR_SHAPE, f"slice: {n_in} indices for a {self.ndim}-d array")
        v = self.v.copy()
        axis = 0  # axis in the array being rewritten
        new_axes = []
        out_axis = 0
        for e in info:
            if isinstance(e, _NewAxis):
                new_axes.append(out_axis)
                out_axis += 1
            elif isinstance(e, Slice):
                check(lib().ndb_view_slice_axis(C.byref(v), axis, e.start, 0 if e.end is None else e.end,
                                                0 if e.end is None else 1, e.step))
                axis += 1
                out_axis += 1
            else:
                idx = e + v.dim[axis] if e < 0 else e
                check(lib().ndb_view_index_axis(C.byref(v), axis, idx))
        out = self._derive(v)
        for ax in new_axes:
            out = out.insert_axis(ax)
        return out

    def index_axis(self, axis: int, index: int) -> "Array":
        v = self.v.copy()
        check(lib().ndb_view_index_axis(C.byref(v), int(axis), int(index)))
        return self._derive(v)

    def insert_axis(self, axis: int) -> "Array":
        axis = int(axis)
        if not 0 <= axis <= self.ndim or self.ndim + 1 > _ffi.MAX_NDIM:
            raise _ffi.NdbError(_ffi.ERR_OOB, "insert_axis: axis out of bounds")
        v = self.v.copy()
        for i in range(self.ndim, axis, -1):
            v.dim[i] = v.dim[i - 1]
            v.stride[i] = v.stride[i - 1]
        v.dim[axis] = 1
        v.stride[axis] = 1
        v.ndim = self.ndim + 1
        return self._derive(v)

    def permuted_axes(self, perm: Iterable[int]) -> "Array":
        perm = list(perm)
        if len(perm) != self.ndim:
            raise _ffi.NdbError(_ffi.ERR_INVALID, "permuted_axes: wrong number of axes")
        out = View()
        arr = (C.c_int * max(len(perm), 1))(*perm)
        check(lib().ndb_view_permute(C.byref(self.v), arr, C.byref(out)))
        return self._derive(out)

    def reversed_axes(self) -> "Array":
        return self.permuted_axes(range(self.ndim - 1, -1, -1))

    def t(self) -> "Array":
        return self.reversed_axes()

    def swap_axes(self, a: int, b: int) -> "Array":
        perm = list(range(self.ndim))
        perm[a], perm[b] = perm[b], perm[a]
        return self.permuted_axes(perm)

    def invert_axis(self, axis: int) -> "Array":
        v = self.v.copy()
        check(lib().ndb_view_invert_axis(C.byref(v), int(axis)))
        return self._derive(v)

    def broadcast(self, shape) -> Optional["Array"]:
        """ArrayRef::broadcast (src/impl_methods.rs:2371): None if the shapes are incompatible."""
        shape = _as_shape(shape)
        out = View()
        arr = (C.c_int64 * max(len(shape), 1))(*shape)
        st = lib().ndb_view_broadcast(C.byref(self.v), len(shape), arr, C.byref(out))
        if st in (_ffi.ERR_SHAPE, _ffi.ERR_OVERFLOW):
            return None
        check(st)
        return self._derive(out)

    def merge_axes(self, take: int, into: int) -> bool:
        """merge_axes (src/impl_methods.rs:2724): in place; True if the layout allowed it."""
        r = lib().ndb_view_merge_axes(C.byref(self.v), int(take), int(into))
        if r < 0:
            raise _ffi.NdbError(-r, lib().ndb_last_error().decode())
        return bool(r)

    def collapse_axis(self, axis: int, index: int) -> None:
        """collapse_axis (src/impl_methods.rs:1075): in place; the axis keeps length 1."""
        check(lib().ndb_view_collapse_axis(C.byref(self.v), int(axis), int(index)))

    def split_at(self, axis: int, index: int):
        l, r = View(), View()
        check(lib().ndb_view_split_at(C.byref(self.v), int(axis), int(index), C.byref(l), C.byref(r)))
        return self._derive(l), self._derive(r)

    # -- host transfer
    def to_owned(self, order: str | None = None) -> "Array":
        if order is None:
            order = "F" if lib().ndb_zip_prefer_f(1, C.byref(self.v)) else "C"
        out = Array._alloc(self.shape, self.dtype, order, self.ctx)
        _zip_call(self.ctx, Op.COPY, out, [self])
        return out

    def to_numpy(self) -> np.ndarray:
        """Download in logical order."""
        src = self
        if not (self.strides == tuple(_c_strides(self.shape))):
            src = self.to_owned("C")
        host = np.empty(self.shape, dtype=self.dtype)
        if host.size:
            check(lib().ndb_d2h(self.ctx.handle, C.c_void_p(host.ctypes.data), C.c_void_p(src.v.ptr), host.nbytes))
        self.ctx.sync()
        return host

    # -- elementwise (src/impl_methods.rs:2796-2896, src/impl_ops.rs)
    def assign(self, rhs: "Array") -> None:
        rhs = _broadcast_to(rhs, self.shape, "assign")
        _zip_call(self.ctx, Op.COPY, self, [rhs])

    def fill(self, x) -> None:
        _zip_call(self.ctx, Op.FILL, self, [], [x])

    def zip_mut_with(self, rhs: "Array", op: int) -> None:
        """self = op(self, rhs), rhs broadcast to self's shape (impl_methods.rs:2882)."""
        rhs = _broadcast_to(rhs, self.shape, "zip_mut_with")
        _zip_call(self.ctx, op, self, [self, rhs])

    def scaled_add(self, alpha, rhs: "Array") -> None:
        """self += alpha * rhs (src/linalg/impl_linalg.rs:399-405)."""
        rhs = _broadcast_to(rhs, self.shape, "scaled_add")
        _zip_call(self.ctx, Op.AXPY, self, [self, rhs], [alpha])

    def mapv(self, op: int, scalars=()) -> "Array":
        out = Array._alloc(self.shape, self.dtype, "F" if lib().ndb_zip_prefer_f(1, C.byref(self.v)) else "C", self.ctx)
        _zip_call(self.ctx, op, out, [self], scalars)
        return out

    def mapv_cast(self, dtype) -> "Array":
        """`a.mapv(|x| x as B)`: element-type conversion with Rust's `as` semantics."""
        out = Array._alloc(self.shape, dtype, "F" if lib().ndb_zip_prefer_f(1, C.byref(self.v)) else "C", self.ctx)
        check(lib().ndb_cast(self.ctx.handle, C.byref(out.v), C.byref(self.v)))
        return out

    astype = mapv_cast

    def mapv_inplace(self, op: int, scalars=()) -> None:
        _zip_call(self.ctx, op, self, [self], scalars)

    # -- float maths maps (src/numeric/impl_float_maths.rs:80-160): `a.exp()`, `a.powi(3)`, `a.clamp(lo, hi)`, ...
    def map_math(self, fn: str, param=0.0) -> "Array":
        return self.mapv(Op.MATH, [MATH_FN[fn], param])

    def abs(self) -> "Array":
        return self.mapv(Op.ABS)

    def sqrt(self) -> "Array":
        return self.mapv(Op.SQRT)

    def clamp(self, lo, hi) -> "Array":
        return self.mapv(Op.CLAMP, [lo, hi])

    def is_all_nan(self) -> bool:
        return self.len() == 0 or self.map_math("is_nan")._reduce_all(Red.MIN) != 0

    def is_any_nan(self) -> bool:
        return self.len() > 0 and self.map_math("is_nan")._reduce_all(Red.MAX) != 0

    def is_all_infinite(self) -> bool:
        return self.len() == 0 or self.map_math("is_infinite")._reduce_all(Red.MIN) != 0

    def is_any_infinite(self) -> bool:
        return self.len() > 0 and self.map_math("is_infinite")._reduce_all(Red.MAX) != 0

    def _binary(self, other, op, op_s, rop_s=None, reflected=False):
        if isinstance(other, Array):
            lhs, rhs = (other, self) if reflected else (self, other)
            shape = _co_broadcast(lhs.shape, rhs.shape)
            l2, r2 = _broadcast_to(lhs, shape, "binary op"), _broadcast_to(rhs, shape, "binary op")
            return Zip.from_(l2).and_(r2).map_collect(op)
        if reflected:
            if rop_s is None:  # commutative
                return self.mapv(op_s, [other])
            return self.mapv(rop_s, [other])
        return self.mapv(op_s, [other])

    def __add__(self, o): return self._binary(o, Op.ADD, Op.ADD_S)
    def __radd__(self, o): return self._binary(o, Op.ADD, Op.ADD_S, None, True)
    def __sub__(self, o): return self._binary(o, Op.SUB, Op.SUB_S)
    def __rsub__(self, o): return self._binary(o, Op.SUB, Op.SUB_S, Op.RSUB_S, True)
    def __mul__(self, o): return self._binary(o, Op.MUL, Op.MUL_S)
    def __rmul__(self, o): return self._binary(o, Op.MUL, Op.MUL_S, None, True)
    def __truediv__(self, o): return self._binary(o, Op.DIV, Op.DIV_S)
    def __rtruediv__(self, o): return self._binary(o, Op.DIV, Op.DIV_S, Op.RDIV_S, True)
    def __mod__(self, o): return self._binary(o, Op.REM, Op.REM_S)
    def __neg__(self): return self.mapv(Op.NEG)

    def _inplace(self, other, op, op_s):
        if isinstance(other, Array):
            self.zip_mut_with(other, op)
        else:
            self.mapv_inplace(op_s, [other])
        return self

    def __iadd__(self, o): return self._inplace(o, Op.ADD, Op.ADD_S)
    def __isub__(self, o): return self._inplace(o, Op.SUB, Op.SUB_S)
    def __imul__(self, o): return self._inplace(o, Op.MUL, Op.MUL_S)
    def __itruediv__(self, o): return self._inplace(o, Op.DIV, Op.DIV_S)
    def __imod__(self, o): return self._inplace(o, Op.REM, Op.REM_S)

    # -- reductions (src/numeric/impl_numeric.rs)
    def _reduce_all(self, redop):
        out = np.zeros((), dtype=self.dtype)
        check(lib().ndb_reduce_all(self.ctx.handle, redop, C.byref(self.v), C.c_void_p(out.ctypes.data)))
        return out[()]

    def sum(self):
        return self._reduce_all(Red.SUM)

    def sum_to_device(self, out: "Array") -> None:
        """sum() whose scalar stays on the device (0-d `out`), asynchronous on the context stream: the
        multi-GPU global sum follows it with an all-reduce on the same stream."""
        check(lib().ndb_reduce_all_dev(self.ctx.handle, Red.SUM, C.byref(self.v), C.c_void_p(out.v.ptr)))

    def product(self):
        return self._reduce_all(Red.PROD)

    def mean(self):
        n = self.len()
        if n == 0:
            return None
        s = self.sum()
        if self.dtype.kind == "f":
            return s / self.dtype.type(n)
        q = abs(int(s)) // n  # Rust integer division truncates toward zero
        return self.dtype.type(-q if int(s) < 0 else q)

    def var(self, ddof):
        """ArrayRef::var (src/numeric/impl_numeric.rs:180): Welford over every element / (n - ddof)."""
        out = np.zeros((), dtype=self.dtype)
        check(lib().ndb_var(self.ctx.handle, C.byref(self.v), float(ddof), 0, C.c_void_p(out.ctypes.data)))
        return out[()]

    def std(self, ddof):
        out = np.zeros((), dtype=self.dtype)
        check(lib().ndb_var(self.ctx.handle, C.byref(self.v), float(ddof), 1, C.c_void_p(out.ctypes.data)))
        return out[()]

    def _reduce_axis(self, redop, axis, ddof=0.0) -> "Array":
        axis = int(axis)
        if not 0 <= axis < self.ndim:
            raise _ffi.NdbError(_ffi.ERR_OOB, f"axis {axis} is out of bounds for array of dimension {self.ndim}")
        shape = self.shape[:axis] + self.shape[axis + 1:]
        out = Array._alloc(shape, self.dtype, "C", self.ctx)
        check(lib().ndb_reduce_axis(self.ctx.handle, redop, axis, C.byref(self.v), C.byref(out.v), float(ddof)))
        return out

    def sum_axis(self, axis) -> "Array":
        return self._reduce_axis(Red.SUM, axis)

    def product_axis(self, axis) -> "Array":
        return self._reduce_axis(Red.PROD, axis)

    def mean_axis(self, axis) -> Optional["Array"]:
        if self.shape[int(axis)] == 0:
            return None
        return self._reduce_axis(Red.MEAN, axis)

    def var_axis(self, axis, ddof) -> "Array":
        return self._reduce_axis(Red.VAR, axis, ddof)

    def std_axis(self, axis, ddof) -> "Array":
        return self._reduce_axis(Red.STD, axis, ddof)

    def fold_axis(self, axis, init, fold: str) -> "Array":
        """fold_axis(axis, init, closure) for the associative closures: 'add' (wrapping for ints), 'mul',
        'min', 'max' (src/impl_methods.rs:3130).  `init` must be the closure's identity or, for add/mul,
        any value (it is combined once, which is the same thing for an associative, commutative fold)."""
        red = {"add": Red.SUM, "mul": Red.PROD, "min": Red.MIN, "max": Red.MAX}[fold]
        if self.shape[int(axis)] == 0:
            return Array.from_elem(self.shape[:int(axis)] + self.shape[int(axis) + 1:], init, self.dtype, ctx=self.ctx)
        out = self._reduce_axis(red, axis)
        if fold == "add" and init != 0:
            out.mapv_inplace(Op.ADD_S, [init])
        elif fold == "mul" and init != 1:
            out.mapv_inplace(Op.MUL_S, [init])
        elif fold in ("min", "max"):
            ini = Array.from_elem(out.shape, init, self.dtype, ctx=self.ctx)
            _zip_call(self.ctx, Op.MIN if fold == "min" else Op.MAX, out, [ini, out])
        return out

    def diff(self, n: int, axis) -> "Array":
        """diff (src/numeric/impl_numeric.rs:495-532): n-th forward difference along `axis`, one Zip per order."""
        axis = int(axis)
        if n == 0:
            return self.to_owned()
        if not 0 <= axis < self.ndim:
            raise _ffi.NdbError(_ffi.ERR_OOB, f"The array has only ndim {self.ndim}, but `axis` Axis({axis}) is given.")
        if not n < self.shape[axis]:
            raise _ffi.NdbError(_ffi.ERR_INVALID, f"The array must have length at least `n+1`=={n + 1} in the direction of "
                                f"`axis`. It has length {self.shape[axis]}")
        inp = self
        for _ in range(n):
            head = inp.slice_axis(axis, Slice(0, -1, 1))
            tail = inp.slice_axis(axis, Slice(1, None, 1))
            out = Array._alloc(head.shape, self.dtype, "C", self.ctx)
            _zip_call(self.ctx, Op.SUB, out, [tail, head])
            inp = out
        return inp

    def accumulate_axis_inplace(self, axis, op: int) -> None:
        """accumulate_axis_inplace (src/impl_methods.rs:3243): a[i] = op(a[i-1], a[i]) along `axis` (op = Op.ADD or
        Op.MUL), one scan launch."""
        axis = int(axis)
        if not 0 <= axis < self.ndim:
            raise _ffi.NdbError(_ffi.ERR_OOB, f"axis {axis} is out of bounds for array of dimension {self.ndim}")
        check(lib().ndb_accumulate_axis(self.ctx.handle, op, axis, C.byref(self.v)))

    def cumprod(self, axis) -> "Array":
        """cumprod (src/numeric/impl_numeric.rs:125)."""
        out = self.to_owned("C")
        out.accumulate_axis_inplace(axis, Op.MUL)
        return out

    def dot(self, rhs: "Array"):
        """Dot::dot (src/linalg/impl_linalg.rs:160-386): 1-d·1-d -> scalar, 2-d·2-d -> matrix, 2-d·1-d, 1-d·2-d."""
        from .linalg import dot as _dot
        return _dot(self, rhs)

    def __repr__(self):
        return f"Array(shape={self.shape}, strides={self.strides}, dtype={self.dtype})"


def NdbErrorUnsupported(msg):
    return _ffi.NdbError(_ffi.ERR_UNSUPPORTED, msg)


def _co_broadcast(a: tuple, b: tuple) -> tuple:
    """co_broadcast (src/dimension/broadcast.rs:9-35)."""
    n = max(len(a), len(b))
    out = [1] * n
    for i in range(1, n + 1):
        da = a[-i] if i <= len(a) else 1
        db = b[-i] if i <= len(b) else 1
        if da == db:
            out[-i] = da
        elif da == 1:
            out[-i] = db
        elif db == 1:
            out[-i] = da
        else:
            raise _ffi.ShapeError(_ffi.ERR_SHAPE, f"ShapeError/IncompatibleShape: incompatible shapes {a} and {b}")
    return tuple(out)


def _broadcast_to(a: Array, shape: tuple, what: str) -> Array:
    if a.shape == tuple(shape):
        return a
    b = a.broadcast(shape)
    if b is None:
        # ndarray: broadcast_unwrap panics (src/lib.rs:1810)
        raise _ffi.ShapeError(_ffi.ERR_SHAPE,
                              f"ndarray: could not broadcast array from shape: {list(a.shape)} to: {list(shape)}")
    return b


def _zip_call(ctx: Context, op: int, out: Array, inputs: Sequence[Array], scalars=()) -> None:
    n = 1 + len(inputs)
    parts = (View * n)(out.v, *[x.v for x in inputs])
    sc = None
    if len(scalars):
        sc = np.asarray(scalars, dtype=out.dtype)
    check(lib().ndb_zip(ctx.handle, op, C.c_void_p(sc.ctypes.data) if sc is not None else None, n, parts))


# closures the reference passes to Zip, by name: (op, index of the written producer, input producer indices)
_CLOSURES = {
    # *a = b
    "a=b": (Op.COPY, 0, (1,)),
    "a=b+c": (Op.ADD, 0, (1, 2)),
    "a=b-c": (Op.SUB, 0, (1, 2)),
    "a=b*c": (Op.MUL, 0, (1, 2)),
    "a=b/c": (Op.DIV, 0, (1, 2)),
    "a+=b": (Op.ADD, 0, (0, 1)),
    "a-=b": (Op.SUB, 0, (0, 1)),
    "a*=b": (Op.MUL, 0, (0, 1)),
    "a/=b": (Op.DIV, 0, (0, 1)),
    # BASELINE config 1: Zip::from(&mut a).and(&b).and(&c).for_each(|a, &b, &c| *a = b * c + *a)
    "a=b*c+a": (Op.MUL_ADD, 0, (1, 2, 0)),
    "a=b*c+d": (Op.MUL_ADD, 0, (1, 2, 3)),
}


class Zip:
    """Lock-step traversal of 1..9 producers of equal shape (src/zip/mod.rs:193)."""

    def __init__(self, parts: list):
        self.parts = parts

    @staticmethod
    def from_(p: Array) -> "Zip":
        return Zip([p])

    def and_(self, p: Array) -> "Zip":
        if p.shape != self.parts[0].shape:
            # zip_dimension_check panics (src/zip/mod.rs:247-258)
            raise _ffi.ShapeError(_ffi.ERR_SHAPE, "Zip: Producer dimension mismatch, expected: "
                                  f"{list(self.parts[0].shape)}, got: {list(p.shape)}")
        if len(self.parts) >= _ffi.MAX_PARTS:
            raise NdbErrorUnsupported("Zip: at most 9 producers")
        return Zip(self.parts + [p])

    def and_broadcast(self, p: Array) -> "Zip":
        return Zip(self.parts + [_broadcast_to(p, self.parts[0].shape, "and_broadcast")])

    def _ctx(self):
        return self.parts[0].ctx

    def for_each(self, closure, scalars=()) -> None:
        """Apply a closure from the vocabulary.  `closure` is a name from _CLOSURES (which producer is
        written and which are read is part of the name, as it is part of a Rust closure), or an Op code,
        in which case the FIRST producer is written and the others are read in order."""
        if isinstance(closure, str):
            if closure.startswith("a=sum"):  # "a=sum(rest)": n-ary add (tests/azip.rs:492)
                op, out_i, ins = Op.ADD_N, 0, tuple(range(1, len(self.parts)))
            else:
                op, out_i, ins = _CLOSURES[closure]
        else:
            op, out_i, ins = closure, 0, tuple(range(1, len(self.parts)))
        _zip_call(self._ctx(), op, self.parts[out_i], [self.parts[i] for i in ins], scalars)

    par_for_each = for_each  # the GPU grid is the rayon pool (src/parallel/impl_par_methods.rs:64)

    def for_each_fold(self, closure, fold: str = "sum", scalars=(), out: "Array | None" = None, across_ranks: bool = False):
        """Zip::fold whose closure also writes its first producer (src/zip/mod.rs:622), e.g.
        `fold(0., |acc, a, &b, &c| { *a = b*c + *a; acc + *a })` = for_each_fold("a=b*c+a", "sum"): one fused pass when the
        producers are dense.  Returns the folded value (host), or leaves it in the 0-d device array `out` (asynchronous)."""
        redop = {"sum": Red.SUM, "prod": Red.PROD, "min": Red.MIN, "max": Red.MAX}[fold]
        if isinstance(closure, str):
            op, out_i, ins = _CLOSURES[closure]
        else:
            op, out_i, ins = closure, 0, tuple(range(1, len(self.parts)))
        parts_l = [self.parts[out_i]] + [self.parts[i] for i in ins]
        n = len(parts_l)
        views = (View * n)(*[p.v for p in parts_l])
        dt = self.parts[0].dtype
        sc = np.asarray(scalars, dtype=dt) if len(scalars) else None
        dev = out if out is not None else Array.zeros((), dt, ctx=self._ctx())
        check(lib().ndb_zip_fold_write(self._ctx().handle, op, C.c_void_p(sc.ctypes.data) if sc is not None else None, n, views, redop,
                                       C.c_void_p(dev.v.ptr), int(across_ranks)))
        return None if out is not None else dev.to_numpy()[()]

    def map_collect(self, op: int, scalars=()) -> Array:
        """Zip::map_collect (src/zip/mod.rs:764-789): fresh output, F-order iff prefer_f()."""
        n = len(self.parts)
        views = (View * n)(*[p.v for p in self.parts])
        order = "F" if lib().ndb_zip_prefer_f(n, views) else "C"
        out = Array._alloc(self.parts[0].shape, self.parts[0].dtype, order, self._ctx())
        _zip_call(self._ctx(), op, out, self.parts, scalars)
        return out

    par_map_collect = map_collect

    def map_assign_into(self, into: Array, op: int, scalars=()) -> None:
        if into.shape != self.parts[0].shape:
            raise _ffi.ShapeError(_ffi.ERR_SHAPE, "Zip: Producer dimension mismatch, expected: "
                                  f"{list(self.parts[0].shape)}, got: {list(into.shape)}")
        _zip_call(self._ctx(), op, into, self.parts, scalars)

    par_map_assign_into = map_assign_into

    def _pred(self, fn, op: int, scalars=()) -> bool:
        n = len(self.parts)
        views = (View * n)(*[p.v for p in self.parts])
        sc = np.asarray(scalars, dtype=self.parts[0].dtype) if len(scalars) else None
        res = C.c_int()
        check(fn(self._ctx().handle, op, C.c_void_p(sc.ctypes.data) if sc is not None else None, n, views, C.byref(res)))
        return bool(res.value)

    def all(self, op: int, scalars=()) -> bool:
        """Zip::all (src/zip/mod.rs:658) with a predicate op (Op.EQ .. Op.GE) over ALL producers (none is written)."""
        return self._pred(lib().ndb_zip_all, op, scalars)

    def any(self, op: int, scalars=()) -> bool:
        """Zip::any (src/zip/mod.rs:685)."""
        return self._pred(lib().ndb_zip_any, op, scalars)

    def fold_sum(self, op: int, scalars=()):
        """Zip::fold(0, |acc, ...| acc + op(...)) / par_fold (src/zip/mod.rs:622, impl_par_methods.rs:173)."""
        n = len(self.parts)
        views = (View * n)(*[p.v for p in self.parts])
        sc = np.asarray(scalars, dtype=self.parts[0].dtype) if len(scalars) else None
        out = np.zeros((), dtype=self.parts[0].dtype)
        check(lib().ndb_zip_fold_sum(self._ctx().handle, op, C.c_void_p(sc.ctypes.data) if sc is not None else None, n, views,
                                     C.c_void_p(out.ctypes.data)))
        return out[()]

    par_fold_sum = fold_sum

    def fold(self, init, op: int, fold: str = "sum", scalars=()):
        """Zip::fold(init, |acc, ...| acc <fold> op(...)) for an associative accumulator (src/zip/mod.rs:622):
        fold in {"sum", "prod", "min", "max"}.  An empty Zip returns `init`."""
        redop = {"sum": Red.SUM, "prod": Red.PROD, "min": Red.MIN, "max": Red.MAX}[fold]
        n = len(self.parts)
        dt = self.parts[0].dtype
        views = (View * n)(*[p.v for p in self.parts])
        sc = np.asarray(scalars, dtype=dt) if len(scalars) else None
        ini = np.asarray(init, dtype=dt) if init is not None else None
        out = np.zeros((), dtype=dt)
        check(lib().ndb_zip_fold(self._ctx().handle, op, C.c_void_p(sc.ctypes.data) if sc is not None else None, n, views, redop,
                                 C.c_void_p(ini.ctypes.data) if ini is not None else None, C.c_void_p(out.ctypes.data)))
        return out[()]

    par_fold = fold

    def fold_while(self, init, op: int, fold: str, stop_op: int, stop_value, scalars=()):
        """Zip::fold_while (src/zip/mod.rs:640): fold op(inputs) in traversal order until the first value with
        `value <stop_op> stop_value`; returns (accumulator, done) where done is True for FoldWhile::Done."""
        redop = {"sum": Red.SUM, "prod": Red.PROD, "min": Red.MIN, "max": Red.MAX}[fold]
        n = len(self.parts)
        dt = self.parts[0].dtype
        views = (View * n)(*[p.v for p in self.parts])
        sc = np.asarray(scalars, dtype=dt) if len(scalars) else None
        ini = np.asarray(init, dtype=dt) if init is not None else None
        sv = np.asarray(stop_value, dtype=dt)
        out = np.zeros((), dtype=dt)
        done = C.c_int()
        check(lib().ndb_zip_fold_while(self._ctx().handle, op, C.c_void_p(sc.ctypes.data) if sc is not None else None, n, views, redop,
                                       C.c_void_p(ini.ctypes.data) if ini is not None else None, stop_op, C.c_void_p(sv.ctypes.data),
                                       C.c_void_p(out.ctypes.data), C.byref(done)))
        return out[()], bool(done.value)

    @staticmethod
    def indexed(p: "Array") -> "Zip":
        """Zip::indexed (src/zip/mod.rs:237): the index along each axis rides along as extra producers (broadcast
        aranges — the Indices producer of src/indexes.rs as stride-0/1 views), after `p`: parts = [p, i0, i1, ...]."""
        z = Zip([p])
        for ix in indices(p.shape, p.dtype, p.ctx):
            z = z.and_(ix)
        return z

    def plan(self, op: int = Op.ADD_N):
        """The kernel class the planner picks for (first producer written, rest read)."""
        n = len(self.parts)
        views = (View * n)(*[p.v for p in self.parts])
        pl = _ffi.ZipPlan()
        check(lib().ndb_zip_plan(op, n, views, C.byref(pl)))
        return pl


def indices(shape, dtype=np.int64, ctx=None) -> list:
    """One producer per axis holding that axis' index, broadcast to `shape` (ndarray::indices, src/indexes.rs)."""
    shape = _as_shape(shape)
    out = []
    for k, d in enumerate(shape):
        ar = Array.from_numpy(np.arange(d, dtype=dtype), ctx)
        v = View()
        v.ptr, v.ndim, v.dtype = ar.v.ptr, len(shape), ar.v.dtype
        for i, di in enumerate(shape):
            v.dim[i] = di
            v.stride[i] = 1 if (i == k and di > 1) else 0
        out.append(ar._derive(v))
    return out


def azip(closure, *producers, scalars=()):
    """`azip!((a in &mut a, &b in &b, ...) closure)` (src/zip/zipmacro.rs:98): sugar for Zip::from(a).and(b)....for_each."""
    z = Zip.from_(producers[0])
    for p in producers[1:]:
        z = z.and_(p)
    z.for_each(closure, scalars)


par_azip = azip  # src/parallel/zipmacro.rs:53


def _install_math_methods():
    unary = [f for f in MATH_FNS if f not in ("powi", "powf", "log", "abs_sub", "hypot")]
    for name in unary:
        setattr(Array, name, (lambda n: lambda self: self.map_math(n))(name))
    for name in ("powi", "powf", "log", "abs_sub", "hypot"):
        setattr(Array, name, (lambda n: lambda self, p: self.map_math(n, p))(name))


_install_math_methods()
