"""GPU parity: reductions (ndb_reduce_all / ndb_reduce_axis / ndb_dot through the C ABI) against the oracle.

Integers: bit-exact (wrapping arithmetic is associative and commutative, so the tree order is irrelevant).
Floats: the summation order differs from ndarray's by design, so results are compared with the oracle's
reference-order value AND a compensated exact sum, within the north-star tolerance
    |got - exact| <= tol * sqrt(n) * max|x|      tol = 1e-5 (f32), 1e-12 (f64)
which is also the bound the reference-order result itself satisfies."""
import math

import numpy as np
import pytest

from .conftest import assert_bit_equal, device_view_of
from .test_zip_gpu import Mirror, rand

pytestmark = pytest.mark.gpu
RED = dict(SUM=0, PROD=1, MIN=2, MAX=3, MEAN=4, VAR=5, STD=6)
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}

ident = lambda a: a
VIEWS = {
    "c2": ((300, 257), ident),
    "f2": ((257, 300), lambda a: a.T),
    "c3": ((13, 40, 70), ident),
    "perm3": ((70, 13, 40), lambda a: a.transpose(1, 2, 0)),
    "neg": ((40, 66), lambda a: a[::-1, ::-1]),
    "strided": ((40, 132), lambda a: a[::2, ::3]),
    "long_lane": ((3, 70000), ident),
    "tall": ((70000, 3), ident),
    "vec1": ((5000,), ident),
    "vec_rev": ((5001,), lambda a: a[::-1]),
    "vec_s2": ((5001,), lambda a: a[1::2]),
    "one": ((1, 1), ident),
    "aligned_wide": ((64, 4096), ident),
    "bcast": ((17,), lambda a: np.broadcast_to(a, (9, 17))),
}


def float_close(got, exact, n, scale, dtype, what):
    tol = TOL[np.dtype(dtype)] * math.sqrt(max(n, 1)) * max(scale, 1e-30)
    err = np.max(np.abs(np.asarray(got, np.float64) - np.asarray(exact, np.float64)))
    assert err <= tol, f"{what}: err {err:g} > tol {tol:g}"


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("name", sorted(VIEWS))
def test_sum_all(nd, oracle, name, dtype):
    shape, fn = VIEWS[name]
    rng = np.random.default_rng(3)
    m = Mirror(nd, rand(rng, shape, dtype))
    hv, dv = m.view(fn)
    got = dv.sum()
    if np.dtype(dtype).kind == "i":
        assert got == oracle.reduce_all(RED["SUM"], hv)
        with np.errstate(over="ignore"):
            assert got == hv.sum(dtype=dtype)
    else:
        exact = oracle.sum_exact(hv)
        float_close(got, exact, hv.size, float(np.abs(hv).max()), dtype, f"sum {name}")
        float_close(oracle.reduce_all(RED["SUM"], hv), exact, hv.size, float(np.abs(hv).max()), dtype, "oracle order")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("name", sorted(k for k in VIEWS if len(VIEWS[k][0]) >= 1))
@pytest.mark.parametrize("redop", ["SUM", "MEAN", "MIN", "MAX", "PROD"])
def test_axis_reductions(nd, oracle, name, dtype, redop):
    shape, fn = VIEWS[name]
    rng = np.random.default_rng(4)
    host = rand(rng, shape, dtype)
    if redop == "PROD" and np.dtype(dtype).kind == "f":
        host = (1.0 + host * 0.01).astype(dtype)  # keep products finite
    m = Mirror(nd, host)
    hv, dv = m.view(fn)
    for axis in range(hv.ndim):
        st, want = oracle.reduce_axis(RED[redop], hv, axis)
        assert st == 0
        got = dv._reduce_axis(RED[redop], axis).to_numpy()
        n = hv.shape[axis]
        if np.dtype(dtype).kind == "i" or redop in ("MIN", "MAX"):
            assert_bit_equal(got, want, f"{redop} {name} axis {axis}")
        elif redop == "PROD":
            np.testing.assert_allclose(got, want, rtol=TOL[np.dtype(dtype)] * math.sqrt(n) * 4)
        else:
            exact = hv.astype(np.float64).sum(axis=axis) if np.dtype(dtype) == np.float32 else \
                np.array([math.fsum(x) for x in np.moveaxis(hv, axis, -1).reshape(-1, n)]).reshape(want.shape)
            if redop == "MEAN":
                exact = exact / n
            scale = float(np.abs(hv).max()) * (1.0 / n if redop == "MEAN" else 1.0)
            float_close(got, exact, n, scale, dtype, f"{redop} {name} axis {axis}")
            float_close(want, exact, n, scale, dtype, f"oracle {redop} {name} axis {axis}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["c2", "f2", "c3", "perm3", "neg", "strided", "tall", "long_lane", "vec1", "aligned_wide"])
def test_var_std_axis(nd, oracle, name, dtype):
    shape, fn = VIEWS[name]
    rng = np.random.default_rng(5)
    m = Mirror(nd, (rand(rng, shape, dtype) * 3 + 10).astype(dtype))  # a mean that punishes naive sum-of-squares
    hv, dv = m.view(fn)
    for axis in range(hv.ndim):
        n = hv.shape[axis]
        for ddof in (0.0, 1.0, 1.5):
            if ddof > n:
                continue
            st, want = oracle.reduce_axis(RED["VAR"], hv, axis, ddof)
            assert st == 0
            got = dv.var_axis(axis, ddof).to_numpy()
            exact = hv.astype(np.float64).var(axis=axis) * n / (n - ddof) if n - ddof != 0 else want
            scale = float(np.abs(exact).max()) if np.isfinite(exact).all() else 1.0
            tol = TOL[np.dtype(dtype)] * math.sqrt(n) * max(scale, 1e-30) * 16
            fin = np.isfinite(want)
            assert np.array_equal(np.isnan(got), np.isnan(want)) or not fin.all()
            assert np.max(np.abs(got[fin].astype(np.float64) - exact[fin])) <= tol, (name, axis, ddof)
            assert np.max(np.abs(want[fin].astype(np.float64) - exact[fin])) <= tol
            gs = dv.std_axis(axis, ddof).to_numpy()
            np.testing.assert_allclose(gs[fin], np.sqrt(got[fin]), rtol=2e-6 if dtype == np.float32 else 1e-15)


def test_reference_golden_vectors_on_device(nd):
    from ndarray_b200 import Array, Axis
    a = Array.from_numpy(np.array([[1., 2.], [3., 4.]]))
    # tests/numeric.rs:41-54
    assert a.sum_axis(Axis(0)).to_numpy().tolist() == [4., 6.]
    assert a.sum_axis(Axis(1)).to_numpy().tolist() == [3., 7.]
    assert a.product_axis(Axis(0)).to_numpy().tolist() == [3., 8.]
    assert a.product_axis(Axis(1)).to_numpy().tolist() == [2., 12.]
    assert a.mean_axis(Axis(0)).to_numpy().tolist() == [2., 3.]
    assert a.mean_axis(Axis(1)).to_numpy().tolist() == [1.5, 3.5]
    assert a.sum_axis(Axis(1)).sum_axis(Axis(0)).to_numpy()[()] == 10.
    assert a.product_axis(Axis(1)).product_axis(Axis(0)).to_numpy()[()] == 24.
    assert a.sum() == 10. and a.product() == 24.
    # :57-75 empty conventions
    e = Array.ones((2, 0, 3), np.float32)
    assert e.sum() == 0. and e.product() == 1.
    assert Array.ones(0, np.float32).sum_axis(Axis(0)).to_numpy()[()] == 0.
    assert Array.ones(0, np.float32).product_axis(Axis(0)).to_numpy()[()] == 1.
    assert np.array_equal(e.sum_axis(Axis(1)).to_numpy(), np.zeros((2, 3), np.float32))
    assert np.array_equal(e.product_axis(Axis(1)).to_numpy(), np.ones((2, 3), np.float32))
    assert Array.ones(0, np.float32).mean_axis(Axis(0)) is None
    assert e.mean_axis(Axis(1)) is None
    assert e.mean() is None
    # :236-291 var_axis with fractional ddof
    from .test_oracle_golden import VAR_A
    v = Array.from_numpy(VAR_A)
    np.testing.assert_allclose(v.var_axis(Axis(0), 1.5).to_numpy(),
                               [[3.236401e+02, 8.556250e+01, 4.708900e+00, 9.428410e+01],
                                [9.672100e+00, 2.289169e+02, 7.344490e+01, 2.171560e+01],
                                [7.157160e+01, 1.849000e-01, 2.631690e+01, 5.314410e+01]], rtol=0, atol=1e-4)
    np.testing.assert_allclose(v.var_axis(Axis(1), 1.7).to_numpy(),
                               [[0.61676923, 80.81092308, 6.79892308, 0.11789744],
                                [75.19912821, 114.25235897, 48.32405128, 9.03020513]], rtol=0, atol=1e-8)
    np.testing.assert_allclose(v.var_axis(Axis(2), 2.3).to_numpy(),
                               [[79.64552941, 129.09663235, 95.98929412],
                                [109.64952941, 43.28758824, 36.27439706]], rtol=0, atol=1e-8)
    b = Array.from_numpy(np.array([[1.1, 2.3, 4.7]]))
    np.testing.assert_allclose(b.var_axis(Axis(0), 0.).to_numpy(), [0., 0., 0.], atol=1e-12)
    np.testing.assert_allclose(b.var_axis(Axis(1), 0.).to_numpy(), [2.24], atol=1e-12)
    assert Array.zeros((2, 0), np.float64).var_axis(Axis(0), 0.).shape == (0,)
    d = Array.from_numpy(np.array([1.1, 2.7, 3.5, 4.9]))
    np.testing.assert_allclose(d.var_axis(Axis(0), 0.).to_numpy(), 1.8875, atol=1e-12)
    np.testing.assert_allclose(Array.from_numpy(np.array([[100000., 1., 0.01]])).std_axis(Axis(1), 0.).to_numpy(),
                               [47_140.214_021_552_77], atol=1e-6)
    # :355-390 ddof rules
    t = Array.from_numpy(np.array([1., 2., 3.]))
    with pytest.raises(nd.NdbError, match="ddof"):
        t.var_axis(Axis(0), -1.)
    with pytest.raises(nd.NdbError, match="ddof"):
        t.var_axis(Axis(0), 4.)
    z = Array.zeros((2, 3), np.float64).var_axis(Axis(1), float("nan")).to_numpy()
    assert z.shape == (2,) and np.isnan(z).all()
    z = Array.zeros((2, 0), np.float64).var_axis(Axis(1), 0.).to_numpy()
    assert z.shape == (2,) and np.isnan(z).all()
    with pytest.raises(nd.NdbError, match="out of bounds"):
        t.sum_axis(Axis(1))
    # tests/oper.rs:131-164 dot
    x = np.arange(69, dtype=np.float32)
    y = x * np.float32(2.) - np.float32(7.)
    dx, dy = Array.from_numpy(x), Array.from_numpy(y)
    assert abs(dx.dot(dy) - np.float32(197846.)) <= 1e-5 * 197846
    assert Array.from_numpy(x.astype(np.int32)).dot(Array.from_numpy(y.astype(np.int32))) == 197846
    for i in range(1, 8):
        got = dx.slice(nd.s_[i:]).dot(dy.slice(nd.s_[i:]))
        assert abs(got - float(np.dot(x[i:].astype(np.float64), y[i:].astype(np.float64)))) <= 1e-5 * 197846


def test_fold_axis_i64_wrapping(nd, oracle):
    rng = np.random.default_rng(9)
    h = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, size=(513, 300), dtype=np.int64)
    a = nd.Array.from_numpy(h)
    for axis in (0, 1):
        st, want = oracle.reduce_axis(RED["SUM"], h, axis)
        assert_bit_equal(a.fold_axis(axis, 0, "add").to_numpy(), want, "fold add")
        assert_bit_equal(a.fold_axis(axis, 5, "add").to_numpy(), want + np.int64(5), "fold add init")
        assert_bit_equal(a.fold_axis(axis, np.iinfo(np.int64).max, "min").to_numpy(), h.min(axis=axis), "fold min")
        assert_bit_equal(a.fold_axis(axis, np.iinfo(np.int64).min, "max").to_numpy(), h.max(axis=axis), "fold max")


def test_config3_full_size(nd):
    """BASELINE config 3 at full size (16384^2 f64 and i64): column sums / row sums against the
    checksum-of-checksums property sum(sum_axis(0)) == sum(sum_axis(1)) == sum(), exactly for i64 and within
    tolerance for f64, plus sampled lanes against numpy."""
    from ndarray_b200 import Array, Axis
    n = 16384
    rng = np.random.default_rng(10)
    hi = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, size=(n, n), dtype=np.int64)
    a = Array.from_numpy(hi)
    s0, s1 = a.sum_axis(Axis(0)), a.sum_axis(Axis(1))
    total = a.sum()
    assert s0.sum() == total and s1.sum() == total
    with np.errstate(over="ignore"):
        assert np.array_equal(s0.to_numpy()[:64], hi[:, :64].sum(axis=0, dtype=np.int64))
        assert np.array_equal(s1.to_numpy()[:64], hi[:64].sum(axis=1, dtype=np.int64))
        assert total == hi.sum(dtype=np.int64)
    assert np.array_equal(a.fold_axis(Axis(0), np.iinfo(np.int64).max, "min").to_numpy()[:64], hi[:, :64].min(axis=0))
    del a, s0, s1, hi
    hf = rng.standard_normal((n, n))
    f = Array.from_numpy(hf)
    tol = 1e-12 * math.sqrt(n)
    s0, s1 = f.sum_axis(Axis(0)).to_numpy(), f.sum_axis(Axis(1)).to_numpy()
    assert np.max(np.abs(s0[:64] - hf[:, :64].sum(axis=0))) <= tol * 8
    assert np.max(np.abs(s1[:64] - hf[:64].sum(axis=1))) <= tol * 8
    assert abs(s0.sum() - s1.sum()) <= 1e-12 * n * 8
    m0 = f.mean_axis(Axis(0)).to_numpy()
    assert np.max(np.abs(m0 - s0 / n)) <= 1e-15
    v0, v1 = f.var_axis(Axis(0), 1.0).to_numpy(), f.var_axis(Axis(1), 1.0).to_numpy()
    assert np.max(np.abs(v0[:64] - hf[:, :64].var(axis=0, ddof=1))) <= tol
    assert np.max(np.abs(v1[:64] - hf[:64].var(axis=1, ddof=1))) <= tol


def test_var_std_whole_array(nd, oracle):  # tests/numeric.rs:149-233
    from ndarray_b200 import Array
    a = Array.from_numpy(np.array([1., -4.32, 1.14, 0.32]))
    assert abs(a.var(0.) - 5.049875) < 1e-8
    assert abs(a.std(0.) - 2.24719) < 1e-5
    t = Array.from_numpy(np.array([1., 2., 3.]))
    for bad in (-1., 4.):
        with pytest.raises(nd.NdbError, match="ddof"):
            t.var(bad)
        with pytest.raises(nd.NdbError, match="ddof"):
            t.std(bad)
    assert math.isnan(Array.zeros((2, 3), np.float64).var(float("nan")))
    assert math.isnan(Array.zeros(0, np.float64).var(0.0)) and math.isnan(Array.zeros(0, np.float64).std(0.0))
    rng = np.random.default_rng(12)
    for dtype, tol in ((np.float32, 1e-5), (np.float64, 1e-12)):
        h = (rng.standard_normal((300, 257)) * 2 + 5).astype(dtype)
        for fn in (lambda x: x, lambda x: x.T, lambda x: x[::-1, ::2]):
            m = Mirror(nd, h)
            hv, dv = m.view(fn)
            want = float(hv.astype(np.float64).var(ddof=1))
            assert abs(dv.var(1.0) - want) <= tol * math.sqrt(hv.size) * want
            assert abs(dv.std(1.0) - math.sqrt(want)) <= tol * math.sqrt(hv.size) * math.sqrt(want)


def test_diff_cumprod_indexed_golden(nd):
    from ndarray_b200 import Array, Axis, Zip, Op
    d = Array.from_numpy(np.array([1.0, 2.0, 4.0, 7.0]))                     # tests/numeric.rs:412-477
    assert d.diff(1, Axis(0)).to_numpy().tolist() == [1.0, 2.0, 3.0]
    assert d.diff(2, Axis(0)).to_numpy().tolist() == d.diff(1, Axis(0)).diff(1, Axis(0)).to_numpy().tolist()
    assert d.diff(3, Axis(0)).to_numpy().tolist() == d.diff(1, Axis(0)).diff(1, Axis(0)).diff(1, Axis(0)).to_numpy().tolist() == [0.0]
    d2 = Array.from_numpy(np.array([[1.0, 2.0, 4.0, 7.0], [1.0, 3.0, 6.0, 6.0], [1.5, 3.5, 5.5, 5.5]]))
    assert d2.diff(1, Axis(0)).to_numpy().tolist() == [[0.0, 1.0, 2.0, -1.0], [0.5, 0.5, -0.5, -0.5]]
    assert d2.diff(1, Axis(1)).to_numpy().tolist() == [[1.0, 2.0, 3.0], [2.0, 3.0, 0.0], [2.0, 2.0, 0.0]]
    with pytest.raises(nd.NdbError):
        d.diff(10, Axis(0))
    with pytest.raises(nd.NdbError):
        Array.from_numpy(np.array([1, 2, 4, 7])).diff(1, Axis(2))
    a = Array.from_numpy(np.array([1, 2, 3, 4]))                             # tests/numeric.rs:78-137
    assert a.cumprod(Axis(0)).to_numpy().tolist() == [1, 2, 6, 24]
    a2 = Array.from_numpy(np.array([[1, 2], [3, 4]]))
    assert a2.cumprod(Axis(0)).to_numpy().tolist() == [[1, 2], [3, 8]]
    assert a2.cumprod(Axis(1)).to_numpy().tolist() == [[1, 2], [3, 12]]
    a3 = Array.from_numpy(np.array([[[1, 2], [3, 4]], [[5, 6], [7, 8]]]))
    assert a3.cumprod(Axis(0)).to_numpy().tolist() == [[[1, 2], [3, 4]], [[5, 12], [21, 32]]]
    assert a3.cumprod(Axis(1)).to_numpy().tolist() == [[[1, 2], [3, 8]], [[5, 6], [35, 48]]]
    assert a3.cumprod(Axis(2)).to_numpy().tolist() == [[[1, 2], [3, 12]], [[5, 30], [7, 56]]]
    assert Array.zeros((0, 0), np.int32).cumprod(Axis(0)).shape == (0, 0)
    assert Array.from_numpy(np.array([[5]])).cumprod(Axis(1)).to_numpy().tolist() == [[5]]
    with pytest.raises(nd.NdbError, match="out of bounds"):
        a2.cumprod(Axis(2))
    # Zip::indexed (tests/azip.rs:320-457): a[i, j] = i * 10 + j built from the index producers
    out = Array.zeros((5, 7), np.int64)
    z = Zip.indexed(out)
    assert len(z.parts) == 3
    i, j = z.parts[1], z.parts[2]
    ten = Array.from_elem((), 10, np.int64).broadcast((5, 7))
    tmp = Zip.from_(i).and_(ten).map_collect(Op.MUL)
    Zip.from_(out).and_(tmp).and_(j).for_each("a=b+c")
    assert out.to_numpy().tolist() == [[r * 10 + c for c in range(7)] for r in range(5)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
def test_accumulate_axis_scan(nd, dtype):
    """cumprod / running sum along every axis of C-order, transposed and reversed views: integers exact, floats
    bit-identical where a thread walks the axis sequentially (strided axis) and within rounding where a warp scans it."""
    from ndarray_b200 import Array, Op
    rng = np.random.default_rng(19)
    shape = (67, 45, 130)
    if np.dtype(dtype).kind == "i":
        h = rng.integers(-3, 4, size=shape).astype(dtype)
    else:
        h = (1.0 + 0.02 * rng.standard_normal(shape)).astype(dtype)
    for fn in (lambda x: x, lambda x: x.transpose(2, 0, 1), lambda x: x[::-1, :, ::-2]):
        for op, ref in ((Op.MUL, np.cumprod), (Op.ADD, np.cumsum)):
            for axis in range(3):
                m = Mirror(nd, h)
                hv, dv = m.view(fn)
                want_view = ref(hv.astype(np.float64) if np.dtype(dtype).kind == "f" else hv, axis=axis, dtype=None if np.dtype(dtype).kind == "f" else dtype)
                dv.accumulate_axis_inplace(axis, op)
                got = fn(m.download())
                if np.dtype(dtype).kind == "i":
                    assert np.array_equal(got, want_view)
                else:
                    tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(hv.shape[axis])
                    assert np.max(np.abs(got - want_view) / np.maximum(np.abs(want_view), 1e-30)) <= tol
    # a long dense lane: one warp scans 2^20 elements
    long = rng.integers(-1, 2, size=(3, 1 << 20)).astype(np.int64)
    d = Array.from_numpy(long)
    d.accumulate_axis_inplace(1, Op.ADD)
    assert np.array_equal(d.to_numpy(), np.cumsum(long, axis=1))


def test_split_column_reduction_one_launch_equals_two_launches(nd, monkeypatch):
    """sum/var along Axis(0) of a C-order matrix split the reduced axis over blocks.  The single-launch form (the last
    block per column tile merges the partial states in split order) must give the bits of the two-launch form, twice in
    a row (the ticket counters reset themselves)."""
    from ndarray_b200 import Array, Axis
    rng = np.random.default_rng(5)
    for shape, dtype in (((4096, 4096), np.float32), ((3000, 1000), np.float64), ((5000, 96), np.int64)):
        h = (rng.standard_normal(shape) * 100).astype(dtype)
        d = Array.from_numpy(h)
        got = []
        for two in (False, True, False):
            if two:
                monkeypatch.setenv("NDB_RED_TWO_KERNELS", "1")
            else:
                monkeypatch.delenv("NDB_RED_TWO_KERNELS", raising=False)
            res = [d.sum_axis(Axis(0)).to_numpy()]
            if np.dtype(dtype).kind == "f":
                res.append(d.var_axis(Axis(0), 1.0).to_numpy())
            got.append(res)
        monkeypatch.delenv("NDB_RED_TWO_KERNELS", raising=False)
        for a, b, c in zip(*got):
            assert np.array_equal(a, b) and np.array_equal(a, c)


def test_peer_window_world1_and_symbols(nd):
    """The peer-memory exchange with a one-rank world: all-reduce is the identity, all-gather a copy and the fused
    sum + exchange equals the plain sum (the multi-rank protocol runs under torchrun: benchmarks/dist_check.py)."""
    import ctypes as C
    from ndarray_b200 import Array
    from ndarray_b200._ffi import check, lib
    from ndarray_b200.array import DTYPE_CODE, Red
    ctx = nd.Context(0)
    try:
        handle = (C.c_ubyte * 64)()
        check(lib().ndb_peer_window_create(ctx.handle, 0, 1, 1 << 16, handle))
        assert lib().ndb_peer_window_create(ctx.handle, 0, 1, 1 << 16, handle) != 0  # one window per context
        h = np.arange(1000, dtype=np.float64) - 300.0
        d = Array.from_numpy(h, ctx)
        for _ in range(3):
            check(lib().ndb_peer_allreduce(ctx.handle, Red.SUM, DTYPE_CODE[d.dtype], C.c_void_p(d.v.ptr), C.c_void_p(d.v.ptr), 1000))
        assert np.array_equal(d.to_numpy(), h)
        out = Array.uninit((1, 1000), np.float64, ctx=ctx)
        check(lib().ndb_peer_allgather(ctx.handle, DTYPE_CODE[d.dtype], C.c_void_p(d.v.ptr), C.c_void_p(out.v.ptr), 1000))
        assert np.array_equal(out.to_numpy()[0], h)
        s = Array.zeros((), np.float64, ctx=ctx)
        check(lib().ndb_reduce_all_peer(ctx.handle, Red.SUM, C.byref(d.v), C.c_void_p(s.v.ptr)))
        assert s.to_numpy()[()] == d.sum()
        big = Array.zeros((1 << 16,), np.float64, ctx=ctx)  # 512 KiB > the 64 KiB slot
        st = lib().ndb_peer_allreduce(ctx.handle, Red.SUM, 1, C.c_void_p(big.v.ptr), C.c_void_p(big.v.ptr), 1 << 16)
        assert st == 6  # NDB_ERR_OVERFLOW
        check(lib().ndb_peer_window_destroy(ctx.handle))
        assert lib().ndb_peer_allreduce(ctx.handle, Red.SUM, 1, C.c_void_p(d.v.ptr), C.c_void_p(d.v.ptr), 10) == 7  # no window
    finally:
        ctx.close()


def test_multi_gpu_protocol_under_torchrun():
    """Two ranks on two GPUs: sharded reductions with NCCL and with peer-memory exchanges, row-sharded GEMM
    (benchmarks/dist_check.py).  Skipped on a single-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", str(root / "benchmarks" / "dist_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "dist_check ok on 2 GPUs" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("dtype", [np.float32, np.int32, np.float64])
def test_accumulate_axis_vector_scan_with_ragged_tail(nd, dtype):
    """The 128-bit warp scan (dense, 16-byte aligned lanes) on lanes whose length is not a multiple of the vector width:
    rows of 136 elements sliced to 130, so every lane ends in a partial vector and the pad columns must stay untouched."""
    from ndarray_b200 import Array, Op, s_
    rng = np.random.default_rng(23)
    base = (rng.integers(-3, 4, size=(8, 12, 136)) if np.dtype(dtype).kind == "i" else 1.0 + 0.01 * rng.standard_normal((8, 12, 136))).astype(dtype)
    for op, ref in ((Op.ADD, np.cumsum), (Op.MUL, np.cumprod)):
        d = Array.from_numpy(base)
        d.slice(s_[:, :, :130]).accumulate_axis_inplace(2, op)
        got = d.to_numpy()
        assert np.array_equal(got[:, :, 130:], base[:, :, 130:])
        want = ref(base[:, :, :130].astype(np.float64) if np.dtype(dtype).kind == "f" else base[:, :, :130], axis=2)
        if np.dtype(dtype).kind == "i":
            assert np.array_equal(got[:, :, :130], want.astype(dtype))
        else:
            tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(130)
            assert np.max(np.abs(got[:, :, :130] - want) / np.abs(want)) <= tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
def test_accumulate_long_strided_axis_is_chunked(nd, dtype):
    """Few lanes and a long strided axis (cumsum down the rows of a [4096, 96] matrix) take the chunked scan: per-chunk
    totals, carries, local scans.  Integers exactly; floats within the reduction tolerance; a -0.0 first element keeps
    its sign; a transposed view and a stepped axis use the same path."""
    from ndarray_b200 import Array, Op, s_
    rng = np.random.default_rng(29)
    kind = np.dtype(dtype).kind
    h = (rng.integers(-3, 4, size=(4096, 96)) if kind == "i" else 1.0 + 0.001 * rng.standard_normal((4096, 96))).astype(dtype)
    if kind == "f":
        h[0, :4] = -0.0
    for op, ref in ((Op.ADD, np.cumsum), (Op.MUL, np.cumprod)):
        d = Array.from_numpy(h)
        d.accumulate_axis_inplace(0, op)
        got = d.to_numpy()
        want = ref(h.astype(np.float64) if kind == "f" else h, axis=0)
        if kind == "i":
            assert np.array_equal(got, want.astype(dtype))
        else:
            tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(4096)
            assert np.max(np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)) <= tol
            if op == Op.ADD:
                assert np.all(np.signbit(got[0, :4]))
    # stepped rows of a transposed view: lanes along axis 1 of h.T[:, ::3]
    d = Array.from_numpy(h)
    d.t().slice(s_[:, ::3]).accumulate_axis_inplace(1, Op.ADD)
    got = d.to_numpy()
    want = h.copy()
    want[::3, :] = np.cumsum(h[::3, :].astype(np.float64) if kind == "f" else h[::3, :], axis=0).astype(dtype)
    if kind == "i":
        assert np.array_equal(got, want)
    else:
        g64, w64 = got.astype(np.float64), want.astype(np.float64)
        assert np.max(np.abs(g64 - w64) / np.maximum(np.abs(w64), 1e-30)) <= (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(1366)
