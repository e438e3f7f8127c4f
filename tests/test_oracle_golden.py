"""Pin the CPU oracle to ndarray's own golden vectors (CPU only).

Every expected value below is a literal from the reference's tests or doc-tests (file:line cited,
relative to the ndarray 0.17.2 tree).  If the oracle reproduces them, it restates the reference
faithfully on this path; the GPU parity tests then compare the CUDA library with the oracle.
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc

OP = dict(COPY=0, FILL=1, ADD=2, SUB=3, MUL=4, DIV=5, REM=6, MUL_ADD=7, ADD_S=8, SUB_S=9, MUL_S=10, DIV_S=11, REM_S=12,
          RSUB_S=13, RDIV_S=14, NEG=15, AXPY=16, ADD_N=17, ABS=18, SQRT=19, MIN=20, MAX=21, SCALE_ADD=22)
RED = dict(SUM=0, PROD=1, MIN=2, MAX=3, MEAN=4, VAR=5, STD=6)
CORDER, FORDER, CPREFER, FPREFER = 1, 2, 4, 8


@pytest.fixture(scope="module", autouse=True)
def _build():
    orc.build()


def nd_slice(a: np.ndarray, *specs):
    """Apply ndarray-semantics slices (start, end|None, step) per axis through the ORACLE's do_slice and
    return the resulting numpy view (same memory)."""
    v = orc.view(a)
    for ax, (s, e, st) in enumerate(specs):
        assert orc.slice_axis(v, ax, s, e, st) == 0
    return view_to_numpy(v, a)


_KEEP = []


def view_to_numpy(v, keep):
    """numpy array over the memory an OrcView describes."""
    import ctypes
    dtype = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.int64}[v.dtype]
    item = np.dtype(dtype).itemsize
    shape = tuple(v.dim[i] for i in range(v.ndim))
    strides = tuple(v.stride[i] for i in range(v.ndim))
    if any(d == 0 for d in shape):
        return np.zeros(shape, dtype)
    lo = sum((d - 1) * s for d, s in zip(shape, strides) if s < 0)
    hi = sum((d - 1) * s for d, s in zip(shape, strides) if s > 0)
    raw = (ctypes.c_byte * ((hi - lo + 1) * item)).from_address(v.ptr + lo * item)
    flat = np.ctypeslib.as_array(raw).view(dtype)
    _KEEP.append(keep)
    return np.lib.stride_tricks.as_strided(flat[-lo:], shape=shape, strides=tuple(s * item for s in strides))


ALL = (0, None, 1)


# ---- slicing: src/slice.rs:800-806 ----------------------------------------------------------------
def test_negative_step_doc_vectors():
    arr = np.array([0, 1, 2, 3], dtype=np.int64)
    assert nd_slice(arr, (1, 3, -1)).tolist() == [2, 1]
    assert nd_slice(arr, (1, None, -2)).tolist() == [3, 1]
    assert nd_slice(arr, (0, 4, -2)).tolist() == [3, 1]
    assert nd_slice(arr, (0, None, -2)).tolist() == [3, 1]


# tests/array.rs:120-127 (inclusive ranges a..=b are start=a, end=b+1, or None when b == -1)
def test_slice_inclusive_range():
    arr = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.int64)
    assert nd_slice(arr, (1, 2, 1), (1, 3, 1)).tolist() == [[5, 6]]
    assert nd_slice(arr, (1, None, 1), (-2, 3, -1)).tolist() == [[6, 5]]
    assert nd_slice(arr, (0, None, 1), (0, 3, 2)).tolist() == [[1, 3], [4, 6]]


# tests/array.rs:107-116 (edge cases keep shapes, never move the pointer out of bounds)
def test_slice_edge_cases():
    arr = np.zeros((3, 4, 5), dtype=np.int32)
    v = orc.view(arr)
    assert orc.slice_axis(v, 0, 0, 0, -1) == 0
    assert (v.dim[0], v.dim[1], v.dim[2]) == (0, 4, 5)
    assert v.ptr == arr.ctypes.data
    v = orc.raw_view(1000, np.int32, (1, 1), (10, 1))
    assert orc.slice_axis(v, 0, 1, 1, 1) == 0
    assert (v.dim[0], v.dim[1]) == (0, 1) and v.ptr == 1000


def test_slice_out_of_bounds_panics():
    arr = np.zeros(4, dtype=np.int32)
    assert orc.slice_axis(orc.view(arr), 0, 5, None, 1) != 0
    assert orc.slice_axis(orc.view(arr), 0, 0, 5, 1) != 0
    assert orc.slice_axis(orc.view(arr), 0, 0, None, 0) != 0


# ---- layout flags: src/layout/mod.rs:119-242 ---------------------------------------------------------
def test_layout_flags():
    a = np.zeros((5, 5), np.float32)
    b = np.zeros((5, 5), np.float32, order="F")
    assert orc.layout(a) == CORDER | CPREFER
    assert orc.layout(b) == FORDER | FPREFER
    both = CORDER | CPREFER | FORDER | FPREFER
    assert orc.layout(np.zeros((5, 1), np.float32)) == both
    assert orc.layout(np.zeros((1, 5), np.float32, order="F")) == both
    assert orc.layout(np.zeros(5, np.float32)) == both
    assert orc.layout(np.zeros((), np.float32)) == both
    assert orc.layout(nd_slice(a, (0, 1, 1), ALL)) == both         # arow
    assert orc.layout(nd_slice(b, ALL, (0, 1, 1))) == both         # bcol
    assert orc.layout(nd_slice(a, ALL, (0, 1, 1))) == 0            # acol
    assert orc.layout(nd_slice(b, (0, 1, 1), ALL)) == 0            # brow
    # stride_layouts
    assert orc.layout(nd_slice(a, (1, None, 1), ALL)) == CORDER | CPREFER
    assert orc.layout(nd_slice(a, ALL, (1, None, 1))) == CPREFER
    assert orc.layout(nd_slice(b, (1, None, 1), ALL)) == FPREFER
    assert orc.layout(nd_slice(b, ALL, (1, None, 1))) == FORDER | FPREFER
    # skip_layouts
    assert orc.layout(nd_slice(a, (0, None, 2), ALL)) == CPREFER
    assert orc.layout(nd_slice(a, ALL, (0, None, 2))) == 0
    assert orc.layout(nd_slice(b, (0, None, 2), ALL)) == 0
    assert orc.layout(nd_slice(b, ALL, (0, None, 2))) == FPREFER


# tests/dimension.rs:199-225 max_stride_axis; :163-196 (the commented min_stride_axis expectations still
# describe dimension_trait.rs:343-357)
def test_stride_axes():
    assert orc.max_stride_axis(np.zeros(10, np.float32)) == 0
    a = np.zeros((3, 3), np.float32)
    assert orc.max_stride_axis(a) == 0 and orc.max_stride_axis(a.T) == 1
    assert orc.max_stride_axis(np.zeros((1, 3), np.float32)) == 1
    b = np.zeros((2, 3, 4, 5), np.float32)
    assert orc.max_stride_axis(b) == 0
    for ax in range(1, 4):
        assert orc.max_stride_axis(b.swapaxes(0, ax)) == ax
    assert orc.min_stride_axis(np.zeros(10, np.float32)) == 0
    assert orc.min_stride_axis(a) == 1 and orc.min_stride_axis(a.T) == 0
    assert orc.min_stride_axis(b) == 3
    for ax in range(3):
        assert orc.min_stride_axis(b.swapaxes(3, ax)) == ax
    assert orc.min_stride_axis(np.broadcast_to(a, (8, 3, 3))) == 0


# tests/array.rs:2077-2185: contiguity under negative strides / permutations
def test_is_contiguous():
    c = np.arange(24, dtype=np.int32).reshape(2, 3, 4)
    assert orc.is_contiguous(c)
    assert orc.is_contiguous(c.transpose(2, 0, 1))
    assert orc.is_contiguous(c[::-1])
    assert orc.is_contiguous(c[:, ::-1, ::-1])
    assert not orc.is_contiguous(c[:, ::2])
    assert not orc.is_contiguous(c[:, :, 1:])
    assert orc.is_contiguous(np.arange(5, dtype=np.int32)[::-1])
    assert not orc.is_contiguous(np.arange(6, dtype=np.int32)[::2])


# ---- broadcasting: tests/broadcast.rs:4-33 ----------------------------------------------------------------
def test_broadcast_rules():
    def ok(from_shape, to_shape):
        src = np.zeros(from_shape, np.int64)
        out = orc.OrcView()
        v = orc.view(src)
        dims = (orc.C.c_int64 * max(len(to_shape), 1))(*to_shape)
        return orc.lib().orc_broadcast(orc.C.byref(v), len(to_shape), dims, orc.C.byref(out)) == 0

    assert ok((2, 1, 2, 1), (2, 4, 2, 2))
    assert not ok((2, 1), (1,))
    assert not ok((2, 1), ())
    assert ok((2, 1), (2, 1)) and ok((2, 1), (2, 2)) and ok((2, 1), (32, 2, 1))
    assert not ok((2, 1), (32, 1, 2))
    for to in [(), (1,), (3,), (7, 2, 9)]:
        assert ok((), to)


# ---- elementwise: tests/array.rs:1759-1794 (negative strides, non-contiguous memory, broadcasting) ---------
def test_arithmetic_negative_strides_golden():
    buf = np.arange(1, 13, dtype=np.int64)
    s = np.lib.stride_tricks.as_strided(buf, shape=(2, 3, 2), strides=(8, 32, 16))  # .strides((1, 4, 2))
    a = nd_slice(s, (0, None, -1), (0, None, 2), ALL)
    b_full = nd_slice(s, (0, 2, 1), ALL, ALL)[:, -1, :]  # s![..2, -1, ..]
    b = np.broadcast_to(b_full, (2, 2, 2))
    out = np.zeros((2, 2, 2), np.int64)
    assert orc.zip_(OP["ADD"], out, [a, b]) == 0
    assert out.tolist() == [[[11, 15], [20, 24]], [[10, 14], [19, 23]]]
    a2 = nd_slice(s, (0, None, -1), (0, None, 2), ALL)
    b2 = nd_slice(s, ALL, (0, None, 2), ALL)
    assert orc.zip_(OP["ADD"], out, [a2, b2]) == 0
    assert out.tolist() == [[[3, 7], [19, 23]], [[3, 7], [19, 23]]]


# tests/array.rs:1745-1757
def test_arithmetic_broadcast_golden():
    a = np.array([[2], [3], [4]], dtype=np.int64)
    b = np.array([5, 6, 7], dtype=np.int64)
    A, B = np.broadcast_to(a, (3, 3)), np.broadcast_to(b, (3, 3))
    out = np.zeros((3, 3), np.int64)
    orc.zip_(OP["ADD"], out, [A, B]); assert out.tolist() == [[7, 8, 9], [8, 9, 10], [9, 10, 11]]
    orc.zip_(OP["SUB"], out, [A, B]); assert out.tolist() == [[-3, -4, -5], [-2, -3, -4], [-1, -2, -3]]
    orc.zip_(OP["MUL"], out, [A, B]); assert out.tolist() == [[10, 12, 14], [15, 18, 21], [20, 24, 28]]
    orc.zip_(OP["DIV"], out, [B, A]); assert out.tolist() == [[2, 3, 3], [1, 2, 2], [1, 1, 1]]


# tests/azip.rs:481-497
def test_azip9_golden():
    j = np.arange(62, dtype=np.int32)
    ins = [j * k for k in (1, 2, 4, 8, 16, 32, 64, 128)]
    a = np.zeros(62, np.int32)
    assert orc.zip_(OP["ADD_N"], a, ins) == 0
    assert a.tolist() == (j * 255).tolist()


# tests/azip.rs:31-38: copy from a negative-stride, offset slice
def test_azip_negative_stride_copy():
    b = np.fromfunction(lambda i, j: 1.0 / (i + 2 * j + 1), (5, 10), dtype=np.float32).astype(np.float32)
    bv = nd_slice(b, (0, None, -1), (3, None, 1))
    a = np.zeros((5, 7), np.float32)
    assert orc.zip_(OP["COPY"], a, [bv]) == 0
    assert np.array_equal(a, b[::-1, 3:])


# tests/azip.rs:272-297: F-order + permuted-contiguous operands
def test_contiguous_but_not_c_or_f():
    a = np.arange(27, dtype=np.int64).reshape(3, 3, 3).transpose(2, 1, 0)  # reversed_axes: strides [1,3,9]
    b = a.swapaxes(0, 1)
    ans = np.zeros((3, 3, 3), np.int64, order="F")
    assert orc.zip_(OP["ADD"], ans, [a, b]) == 0
    assert ans[0, 1, 2] == a[0, 1, 2] + b[0, 1, 2]
    assert np.array_equal(ans, a + b)


# tests/azip.rs:53-78: map_collect keeps C/F order
def test_map_collect_order_rule():
    b = np.zeros((5, 10), np.float32)
    c = np.zeros((5, 10), np.float32)
    assert not orc.zip_prefer_f([b, c])
    assert orc.zip_prefer_f([b.T, c.T])


def test_integer_division_traps_and_wrap():
    a = np.array([1, -(2 ** 31), 7], np.int32)
    out = np.zeros(3, np.int32)
    assert orc.zip_(OP["DIV"], out, [a, np.array([1, 1, 0], np.int32)]) == 9
    assert orc.zip_(OP["DIV"], out, [a, np.array([1, -1, 1], np.int32)]) == 9
    assert orc.zip_(OP["ADD"], out, [np.array([2 ** 31 - 1] * 3, np.int32), np.array([1, 2, 3], np.int32)]) == 0
    assert out.tolist() == [-(2 ** 31), -(2 ** 31) + 1, -(2 ** 31) + 2]
    assert orc.zip_(OP["DIV"], out, [np.array([-7, 7, -7], np.int32), np.array([2, -2, -2], np.int32)]) == 0
    assert out.tolist() == [-3, -3, 3]  # truncation toward zero


def test_mul_add_is_not_fused():
    # b*c + a with separate roundings differs from fma for these values
    b = np.array([1.0 + 2.0 ** -12], np.float32)
    c = np.array([1.0 + 2.0 ** -12], np.float32)
    a = np.array([-1.0], np.float32)
    out = a.copy()
    assert orc.zip_(OP["MUL_ADD"], out, [b, c, a]) == 0
    two_step = np.float32(np.float32(b[0] * c[0]) + a[0])
    fused = np.float32(math.fma(float(b[0]), float(c[0]), float(a[0]))) if hasattr(math, "fma") else None
    assert out[0] == two_step
    if fused is not None:
        assert fused != two_step


# ---- reductions: tests/numeric.rs ----------------------------------------------------------------------
def test_sum_mean_prod():  # :41-54
    a = np.array([[1., 2.], [3., 4.]])
    assert orc.reduce_axis(RED["SUM"], a, 0)[1].tolist() == [4., 6.]
    assert orc.reduce_axis(RED["SUM"], a, 1)[1].tolist() == [3., 7.]
    assert orc.reduce_axis(RED["PROD"], a, 0)[1].tolist() == [3., 8.]
    assert orc.reduce_axis(RED["PROD"], a, 1)[1].tolist() == [2., 12.]
    assert orc.reduce_axis(RED["MEAN"], a, 0)[1].tolist() == [2., 3.]
    assert orc.reduce_axis(RED["MEAN"], a, 1)[1].tolist() == [1.5, 3.5]
    assert orc.reduce_all(RED["SUM"], a) == 10.
    assert orc.reduce_all(RED["PROD"], a) == 24.


def test_sum_mean_prod_empty():  # :57-75
    e = np.ones((2, 0, 3), np.float32)
    assert orc.reduce_all(RED["SUM"], e) == 0.
    assert orc.reduce_all(RED["PROD"], e) == 1.
    st, out = orc.reduce_axis(RED["SUM"], np.ones(0, np.float32), 0)
    assert st == 0 and out[()] == 0.
    st, out = orc.reduce_axis(RED["PROD"], np.ones(0, np.float32), 0)
    assert st == 0 and out[()] == 1.
    st, out = orc.reduce_axis(RED["SUM"], e, 1)
    assert st == 0 and np.array_equal(out, np.zeros((2, 3), np.float32))
    st, out = orc.reduce_axis(RED["PROD"], e, 1)
    assert st == 0 and np.array_equal(out, np.ones((2, 3), np.float32))
    assert orc.reduce_axis(RED["MEAN"], np.ones(0, np.float32), 0)[0] == 8   # None
    assert orc.reduce_axis(RED["MEAN"], e, 1)[0] == 8


def test_mean_of_floats():  # :24-38
    a = np.array([
        0.99889651, 0.0150731, 0.28492482, 0.83819218, 0.48413156, 0.80710412, 0.41762936,
        0.22879429, 0.43997224, 0.23831807, 0.02416466, 0.6269962, 0.47420614, 0.56275487,
        0.78995021, 0.16060581, 0.64635041, 0.34876609, 0.78543249, 0.19938356, 0.34429457,
        0.88072369, 0.17638164, 0.60819363, 0.250392, 0.69912532, 0.78855523, 0.79140914,
        0.85084218, 0.31839879, 0.63381769, 0.22421048, 0.70760302, 0.99216018, 0.80199153,
        0.19239188, 0.61356023, 0.31505352, 0.06120481, 0.66417377, 0.63608897, 0.84959691,
        0.43599069, 0.77867775, 0.88267754, 0.83003623, 0.67016118, 0.67547638, 0.65220036,
        0.68043427])
    assert abs(orc.reduce_all(RED["SUM"], a) / 50.0 - 0.5475494054) < 1e-9  # approx default epsilon = f64::EPSILON scale


def test_var_scalar():  # :149-188
    a = np.array([1., -4.32, 1.14, 0.32])
    st, v = orc.var(a, 0.)
    assert st == 0 and abs(v - 5.049875) < 1e-8
    assert abs(math.sqrt(v) - 2.24719) < 1e-5
    assert orc.var(np.array([1., 2., 3.]), -1.)[0] == 7  # panics
    assert orc.var(np.array([1., 2., 3.]), 4.)[0] == 7
    st, v = orc.var(np.zeros((2, 3)), float("nan"))
    assert st == 0 and math.isnan(v)
    st, v = orc.var(np.zeros(0), 0.)
    assert st == 0 and math.isnan(v)


VAR_A = np.array([
    [[-9.76, -0.38, 1.59, 6.23], [-8.57, -9.27, 5.76, 6.01], [-9.54, 5.09, 3.21, 6.56]],
    [[8.23, -9.63, 3.76, -3.48], [-5.46, 5.86, -2.81, 1.35], [-1.08, 4.66, 8.34, -0.73]]])


def test_var_axis_golden():  # :236-291
    st, v = orc.reduce_axis(RED["VAR"], VAR_A, 0, 1.5)
    assert st == 0
    np.testing.assert_allclose(v, [[3.236401e+02, 8.556250e+01, 4.708900e+00, 9.428410e+01],
                                   [9.672100e+00, 2.289169e+02, 7.344490e+01, 2.171560e+01],
                                   [7.157160e+01, 1.849000e-01, 2.631690e+01, 5.314410e+01]], rtol=0, atol=1e-4)
    st, v = orc.reduce_axis(RED["VAR"], VAR_A, 1, 1.7)
    np.testing.assert_allclose(v, [[0.61676923, 80.81092308, 6.79892308, 0.11789744],
                                   [75.19912821, 114.25235897, 48.32405128, 9.03020513]], rtol=0, atol=1e-8)
    st, v = orc.reduce_axis(RED["VAR"], VAR_A, 2, 2.3)
    np.testing.assert_allclose(v, [[79.64552941, 129.09663235, 95.98929412],
                                   [109.64952941, 43.28758824, 36.27439706]], rtol=0, atol=1e-8)
    b = np.array([[1.1, 2.3, 4.7]])
    np.testing.assert_allclose(orc.reduce_axis(RED["VAR"], b, 0, 0.)[1], [0., 0., 0.], atol=1e-12)
    np.testing.assert_allclose(orc.reduce_axis(RED["VAR"], b, 1, 0.)[1], [2.24], atol=1e-12)
    st, v = orc.reduce_axis(RED["VAR"], np.zeros((2, 0)), 0, 0.)
    assert st == 0 and v.shape == (0,)
    d = np.array([1.1, 2.7, 3.5, 4.9])
    np.testing.assert_allclose(orc.reduce_axis(RED["VAR"], d, 0, 0.)[1], 1.8875, atol=1e-12)


def test_std_axis_golden():  # :296-352
    a = np.array([
        [[0.22935481, 0.08030619, 0.60827517, 0.73684379], [0.90339851, 0.82859436, 0.64020362, 0.2774583],
         [0.44485313, 0.63316367, 0.11005111, 0.08656246]],
        [[0.28924665, 0.44082454, 0.59837736, 0.41014531], [0.08382316, 0.43259439, 0.1428889, 0.44830176],
         [0.51529756, 0.70111616, 0.20799415, 0.91851457]]])
    np.testing.assert_allclose(orc.reduce_axis(RED["STD"], a, 0, 1.5)[1],
                               [[0.05989184, 0.36051836, 0.00989781, 0.32669847],
                                [0.81957535, 0.39599997, 0.49731472, 0.17084346],
                                [0.07044443, 0.06795249, 0.09794304, 0.83195211]], rtol=0, atol=1e-4)
    np.testing.assert_allclose(orc.reduce_axis(RED["STD"], a, 1, 1.7)[1],
                               [[0.42698655, 0.48139215, 0.36874991, 0.41458724],
                                [0.26769097, 0.18941435, 0.30555015, 0.35118674]], rtol=0, atol=1e-8)
    np.testing.assert_allclose(orc.reduce_axis(RED["STD"], a, 2, 2.3)[1],
                               [[0.41117907, 0.37130425, 0.35332388], [0.16905862, 0.25304841, 0.39978276]],
                               rtol=0, atol=1e-8)
    b = np.array([[100000., 1., 0.01]])
    np.testing.assert_allclose(orc.reduce_axis(RED["STD"], b, 0, 0.)[1], [0., 0., 0.], atol=1e-12)
    np.testing.assert_allclose(orc.reduce_axis(RED["STD"], b, 1, 0.)[1], [47_140.214_021_552_77], atol=1e-6)


def test_var_axis_ddof_rules():  # :355-390
    assert orc.reduce_axis(RED["VAR"], np.array([1., 2., 3.]), 0, -1.)[0] == 7
    assert orc.reduce_axis(RED["VAR"], np.array([1., 2., 3.]), 0, 4.)[0] == 7
    st, v = orc.reduce_axis(RED["VAR"], np.zeros((2, 3)), 1, float("nan"))
    assert st == 0 and v.shape == (2,) and np.isnan(v).all()
    st, v = orc.reduce_axis(RED["VAR"], np.zeros((2, 0)), 1, 0.)
    assert st == 0 and v.shape == (2,) and np.isnan(v).all()


def test_sum_axis_paths_agree_with_definition():
    # fast path (axis == min_stride_axis) and slow path give the same exact integers
    rng = np.random.default_rng(1)
    a = rng.integers(-2 ** 62, 2 ** 62, size=(7, 9, 5), dtype=np.int64)
    for ax in range(3):
        for arr in (a, a.transpose(2, 0, 1), a[::-1, ::2]):
            st, got = orc.reduce_axis(RED["SUM"], arr, ax)
            assert st == 0
            with np.errstate(over="ignore"):
                assert np.array_equal(got, arr.sum(axis=ax, dtype=np.int64))


def test_unrolled_fold_order():
    # 8 interleaved partial sums, combined (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the tail (numeric_util.rs:14-51)
    x = np.array([1e8, 1.0, -1e8, 1.0] * 5 + [3.0, 5.0, 7.0], dtype=np.float32)
    p = [np.float32(0)] * 8
    n8 = (len(x) // 8) * 8
    for i in range(0, n8, 8):
        for j in range(8):
            p[j] = np.float32(p[j] + x[i + j])
    acc = np.float32(0)
    for j in range(4):
        acc = np.float32(acc + np.float32(p[j] + p[j + 4]))
    for v in x[n8:]:
        acc = np.float32(acc + v)
    assert orc.reduce_all(RED["SUM"], x) == acc


# ---- dot / matmul: tests/oper.rs ---------------------------------------------------------------------
def test_dot_product_golden():  # :131-164
    a = np.arange(69, dtype=np.float32)
    b = a * np.float32(2.) - np.float32(7.)
    assert abs(orc.dot(a, b) - np.float32(197846.)) <= 1e-5
    for i in range(1, 8):
        a1, b1 = a[i:], b[i:]
        ref = np.float32(0)
        for x, y in zip(a1, b1):
            ref = np.float32(ref + np.float32(x * y))
        assert abs(orc.dot(a1, b1) - ref) <= 1e-5
    assert orc.dot(a.astype(np.int32), b.astype(np.int32)) == 197846


def _range_mat(m, n, dtype):
    return np.arange(m * n, dtype=dtype).reshape(m, n)


def test_mat_mul_exact_small_ints():  # :321-379: integer-valued f32 products are exact in any order
    a = _range_mat(8, 8, np.float32)
    b = _range_mat(8, 8, np.float32) / np.float32(4.)
    b[:, 0] += np.float32(1.0)
    c = np.zeros((8, 8), np.float32)
    assert orc.gemm(1, a, b, 0, c) == 0
    assert np.array_equal(c, (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
    cf = np.zeros((8, 8), np.float32, order="F")
    assert orc.gemm(1, a.T.copy().T, b, 0, cf) == 0
    assert np.array_equal(cf, c)


def test_gen_mat_mul_sweep():  # :595-635
    alpha, beta = -2.3, math.pi
    sizes = [(4, 4, 4), (8, 8, 8), (17, 15, 16), (4, 17, 3), (17, 3, 22), (19, 18, 2), (16, 17, 15), (15, 16, 17),
             (67, 63, 62)]
    for s1 in (1, 2, -1, -2):
        for s2 in (1, 2, -1, -2):
            for (m, k, n) in sizes:
                a = _range_mat(m, k, np.float64)
                b = _range_mat(k, n, np.float64)
                c = _range_mat(m, n, np.float64)
                av = nd_slice(a, (0, None, s1), (0, None, s2))
                bv = nd_slice(b, (0, None, s2), (0, None, s2))
                cv = nd_slice(c, (0, None, s1), (0, None, s2))
                want = alpha * (av @ bv) + beta * cv
                assert orc.gemm(alpha, av, bv, beta, cv) == 0
                np.testing.assert_allclose(cv, want, rtol=1e-7, atol=1e-12)


def test_gemm_kahan_reference_is_tighter():  # accuracy.rs:143-165
    rng = np.random.default_rng(7)
    a = rng.standard_normal((40, 90)).astype(np.float32)
    b = rng.standard_normal((90, 30)).astype(np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    c0 = np.zeros((40, 30), np.float32)
    c1 = np.zeros((40, 30), np.float32)
    orc.gemm(1, a, b, 0, c0, mode=0)
    orc.gemm(1, a, b, 0, c1, mode=1)
    e0 = np.abs(c0 - exact).max() / np.abs(exact).max()
    e1 = np.abs(c1 - exact).max() / np.abs(exact).max()
    assert e1 <= e0 and e0 < 1e-5


# ---- round 2: the config-5 helpers ------------------------------------------------------------------------------
def test_counter_generator_is_index_addressable():
    """value(seed, i) is a pure function of the flat index: any split of a stream reproduces the whole; known answers
    pin the definition (splitmix64 output i, top 24 / 53 bits) so the device kernel and the oracle cannot drift together."""
    whole = orc.counter_fill(np.empty(1000, np.float32), 0x5EED0001, 0, 3)
    parts = [orc.counter_fill(np.empty(n, np.float32), 0x5EED0001, first, 1) for first, n in ((0, 1), (1, 332), (333, 667))]
    assert np.array_equal(np.concatenate(parts), whole)
    assert whole.min() >= -1.0 and whole.max() < 1.0

    def splitmix(seed, i):
        m = (1 << 64) - 1
        z = (seed + (i + 1) * 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        return z ^ (z >> 31)

    for i in (0, 1, 999):
        h = splitmix(0x5EED0001, i)
        assert whole[i] == np.float32((h >> 40) * 2.0 ** -23 - 1.0)
    f64 = orc.counter_fill(np.empty(4, np.float64), 7, 2 ** 40, 1)
    assert f64[3] == (splitmix(7, 2 ** 40 + 3) >> 11) * 2.0 ** -52 - 1.0
    i64 = orc.counter_fill(np.empty(4, np.int64), 7, 5, 1)
    assert int(i64[2]) == (splitmix(7, 7) ^ (1 << 63)) - (1 << 63)
    i32 = orc.counter_fill(np.empty(4, np.int32), 7, 5, 1)
    assert int(i32[2]) == ((splitmix(7, 7) >> 32) ^ (1 << 31)) - (1 << 31)


def test_parallel_row_sum_idiom():
    """a.axis_iter(Axis(0)).into_par_iter().map(|x| x.sum()).sum() (tests/par_rayon.rs:12-24): integers equal the plain
    sum exactly, floats agree with it within rounding, whatever the thread count; tests/par_rayon.rs' own case (M x N
    array whose row i is filled with i) gives N * sum(i)."""
    m, n = 1024 // 4, 100
    a = np.zeros((m, n), np.float64)
    for i in range(m):
        a[i] = i
    for t in (1, 2, 5):
        assert orc.par_row_sum(a, t) == n * (m * (m - 1) // 2)
    rng = np.random.default_rng(4)
    ai = rng.integers(-2 ** 62, 2 ** 62, (37, 21), dtype=np.int64)
    with np.errstate(over="ignore"):
        assert orc.par_row_sum(ai, 3) == ai.sum(dtype=np.int64) == orc.reduce_all(0, ai)
    af = rng.standard_normal((64, 333)).astype(np.float32)
    assert abs(float(orc.par_row_sum(af, 4)) - af.astype(np.float64).sum()) < 1e-3
    assert abs(float(orc.par_row_sum(af[:, ::3], 4)) - af[:, ::3].astype(np.float64).sum()) < 1e-3


def test_config5_streaming_sums_match_the_materialised_form():
    """orc_c5_sums regenerates a, b, c on the fly: its f64 sum equals the f64 sum of the oracle's materialised
    a = b*c + a, and its ndarray-order f32 sum equals orc_reduce_all (ndarray's own `a.sum()`) bit for bit."""
    n = 8 * 1000 + 5
    a, b, c = (orc.counter_fill(np.empty(n, np.float32), s, 77, 2) for s in (1, 2, 3))
    assert orc.zip_(7, a, [b, c, a]) == 0
    s, sa, f = orc.c5_sums(1, 2, 3, 77, n, 3, True)
    assert abs(s - a.astype(np.float64).sum()) < 1e-9 and abs(sa - np.abs(a.astype(np.float64)).sum()) < 1e-9
    assert f == orc.reduce_all(0, a)


def test_counter_generator_against_the_committed_fixture():
    """tests/golden/counter_rng.json (made by tests/golden/make_counter_rng_golden.py, a pure-Python splitmix64) pins the
    oracle's generator for all four element types and several seeds / offsets, including offsets beyond 2^32."""
    import json
    from pathlib import Path
    fx = json.loads((Path(__file__).parent / "golden" / "counter_rng.json").read_text())
    assert len(fx["cases"]) >= 20
    for c in fx["cases"]:
        n = len(c["i64"])
        f32 = orc.counter_fill(np.empty(n, np.float32), c["seed"], c["first"], 1)
        assert f32.view(np.uint32).tolist() == c["f32_bits"], c
        f64 = orc.counter_fill(np.empty(n, np.float64), c["seed"], c["first"], 1)
        assert [float(x).hex() for x in f64] == c["f64"], c
        assert orc.counter_fill(np.empty(n, np.int32), c["seed"], c["first"], 1).tolist() == c["i32"], c
        assert orc.counter_fill(np.empty(n, np.int64), c["seed"], c["first"], 1).tolist() == c["i64"], c
