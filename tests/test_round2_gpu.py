"""GPU parity for what round 2 added or fixed: the 9-producer host path, Zip::fold on host buffers, Zip::fold with any
associative accumulator, map_assign_into, the index-addressable generator, the scan's first element, and the BASELINE
config sizes (C1 4096^2 bit-exact against the oracle, C4 8192^3 against OpenBLAS + Kahan rows)."""
import ctypes as C

import numpy as np
import pytest

from .conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}


def hview(a):
    from ndarray_b200._ffi import View
    v = View()
    v.ptr = a.__array_interface__["data"][0]
    v.ndim, v.dtype = a.ndim, DT[a.dtype]
    for i in range(a.ndim):
        v.dim[i], v.stride[i] = a.shape[i], a.strides[i] // a.itemsize
    return v


def test_zip_host_nine_producers(nd, oracle):
    """ADD_N with 8 inputs through ndb_zip_host: nine staging slots (slot 8 used to alias other fields)."""
    from ndarray_b200._ffi import View, lib, check
    rng = np.random.default_rng(5)
    ins = [rng.integers(-2 ** 62, 2 ** 62, (33, 70), dtype=np.int64) for _ in range(8)]
    ins[3] = ins[3][::-1]
    ins[5] = np.ascontiguousarray(ins[5].T).T
    out = np.zeros((33, 70), np.int64)
    want = np.zeros_like(out)
    assert oracle.zip_(17, want, ins) == 0
    parts = (View * 9)(hview(out), *[hview(x) for x in ins])
    check(lib().ndb_zip_host(nd.default_context().handle, 17, None, 9, parts))
    assert_bit_equal(out, want, "9-producer host Zip")
    # and again on the device path for good measure
    from ndarray_b200 import Array, Zip
    d = [Array.from_numpy(np.ascontiguousarray(x)) for x in ins]
    o = Array.zeros((33, 70), np.int64)
    z = Zip.from_(o)
    for x in d:
        z = z.and_(x)
    z.for_each("a=sum")
    assert_bit_equal(o.to_numpy(), want, "9-producer device Zip")


@pytest.mark.parametrize("dtype", [np.float32, np.int64])
@pytest.mark.parametrize("rows", [7, 300])  # one slab / many slabs (the 300-row case is 160 MiB per operand)
def test_zip_host_fold_config5_form(nd, oracle, dtype, rows):
    """ndb_zip_host_fold(MUL_ADD, (a; b, c, a), SUM): a = b*c + a in host memory, and the sum of the new a."""
    from ndarray_b200._ffi import View, lib, check
    cols = 1 << 17
    a, b, c = (oracle.counter_fill(np.empty((rows, cols), dtype), seed, 11, 8) for seed in (1, 2, 3))
    want = a.copy()
    assert oracle.zip_(7, want, [b, c, want], threads=8) == 0
    parts = (View * 4)(hview(a), hview(b), hview(c), hview(a))
    res = np.zeros((), dtype)
    check(lib().ndb_zip_host_fold(nd.default_context().handle, 7, None, 4, parts, 0, C.c_void_p(res.ctypes.data)))
    assert_bit_equal(a, want, "host fold: elementwise part")
    if np.dtype(dtype).kind == "i":
        with np.errstate(over="ignore"):
            assert res[()] == want.sum(dtype=np.int64)
    else:
        ref = want.astype(np.float64).sum()
        assert abs(float(res) - ref) <= 1e-5 * np.sqrt(want.size) * np.abs(want).mean()
    # min / max / product folds of a plain copy
    for redop, fn in ((2, np.min), (3, np.max)):
        check(lib().ndb_zip_host_fold(nd.default_context().handle, 0, None, 2, (View * 2)(hview(a), hview(b)), redop, C.c_void_p(res.ctypes.data)))
        assert res[()] == fn(b)
    empty = np.zeros((0, 5), dtype)
    check(lib().ndb_zip_host_fold(nd.default_context().handle, 0, None, 2, (View * 2)(hview(empty), hview(empty)), 0, C.c_void_p(res.ctypes.data)))
    assert res[()] == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_counter_generator_matches_oracle(nd, oracle, dtype):
    """ndb_fill_counter_rng == orc_counter_fill bit for bit, for any shard offset; floats lie in [-1, 1)."""
    from ndarray_b200 import Array
    from ndarray_b200._ffi import lib, check
    ctx = nd.default_context()
    for shape, first in (((1031,), 0), ((37, 129), 2 ** 33 + 5), ((4, 3, 1000), 12345678901234)):
        d = Array.uninit(shape, dtype, ctx=ctx)
        check(lib().ndb_fill_counter_rng(ctx.handle, C.byref(d.v), 0x5EED0001, first))
        want = oracle.counter_fill(np.empty(shape, dtype), 0x5EED0001, first, 4)
        assert_bit_equal(d.to_numpy(), want, f"counter generator {np.dtype(dtype).name} {shape}")
        if np.dtype(dtype).kind == "f":
            assert want.min() >= -1.0 and want.max() < 1.0
    # sharding invariance: two halves filled separately equal the whole
    whole = Array.uninit((64, 256), dtype, ctx=ctx)
    check(lib().ndb_fill_counter_rng(ctx.handle, C.byref(whole.v), 9, 0))
    top, bottom = whole.split_at(0, 20)
    t2, b2 = Array.uninit((20, 256), dtype, ctx=ctx), Array.uninit((44, 256), dtype, ctx=ctx)
    check(lib().ndb_fill_counter_rng(ctx.handle, C.byref(t2.v), 9, 0))
    check(lib().ndb_fill_counter_rng(ctx.handle, C.byref(b2.v), 9, 20 * 256))
    assert_bit_equal(top.to_numpy(), t2.to_numpy())
    assert_bit_equal(bottom.to_numpy(), b2.to_numpy())
    # a non-standard layout is refused
    assert lib().ndb_fill_counter_rng(ctx.handle, C.byref(whole.t().v), 9, 0) == 2


def test_zip_fold_associative_accumulators(nd, oracle):
    """Zip::fold(init, |acc, a, b| acc <fold> op(a, b)) for sum / prod / min / max (src/zip/mod.rs:622), dense and strided."""
    from ndarray_b200 import Array, Zip, Op, s_
    rng = np.random.default_rng(3)
    for dtype in (np.float64, np.int64, np.float32, np.int32):
        kind = np.dtype(dtype).kind
        xh = (rng.integers(-50, 50, (40, 51)) if kind == "i" else rng.uniform(0.9, 1.1, (40, 51))).astype(dtype)
        yh = (rng.integers(-50, 50, (40, 51)) if kind == "i" else rng.uniform(0.9, 1.1, (40, 51))).astype(dtype)
        x, y = Array.from_numpy(xh), Array.from_numpy(yh)
        for views in ((x, y, xh, yh), (x.slice(s_[::-1, ::2]), y.slice(s_[:, ::2]), xh[::-1, ::2], yh[:, ::2])):
            dx, dy, hx, hy = views
            vals = {Op.ADD: hx + hy, Op.MUL: hx * hy, Op.SUB: hx - hy}
            for op, v in vals.items():
                z = Zip.from_(dx).and_(dy)
                assert z.fold(None, op, "min") == v.min()
                assert z.fold(None, op, "max") == v.max()
                assert z.fold(dtype(7), op, "max") == max(dtype(7), v.max())
                assert z.fold(dtype(-100000), op, "min") == -100000
                got = z.fold(dtype(5), op, "sum")
                if kind == "i":
                    assert got == dtype(5) + v.sum(dtype=dtype)
                else:
                    assert abs(got - (5 + v.astype(np.float64).sum())) <= 1e-5 * v.size
        # product of values near 1 (floats) / small ints (wrapping)
        p = Zip.from_(x).and_(y).fold(None, Op.MUL, "prod")
        if kind == "i":
            with np.errstate(over="ignore"):
                assert p == np.prod((xh * yh).astype(dtype), dtype=dtype)
        else:
            assert abs(p / np.prod((xh.astype(np.float64) * yh)) - 1) < 1e-3
    e = Array.zeros((0, 4), np.float32)
    assert Zip.from_(e).and_(e).fold(np.float32(2.5), Op.ADD, "min") == 2.5  # empty: the initial accumulator
    assert Zip.from_(e).and_(e).fold(None, Op.ADD, "prod") == 1.0
    with pytest.raises(Exception):
        Zip.from_(e).and_(e).fold(None, Op.ADD, "max")


def test_map_assign_into(nd, oracle):
    """Zip::map_assign_into (src/zip/mod.rs:796-804): the closure's result goes into a caller-provided producer of any
    layout; a shape mismatch panics with Zip's message."""
    from ndarray_b200 import Array, Zip, Op, s_
    from ndarray_b200 import _ffi
    rng = np.random.default_rng(8)
    xh = rng.standard_normal((30, 44)).astype(np.float32)
    yh = rng.standard_normal((44, 30)).astype(np.float32)
    x, y = Array.from_numpy(xh), Array.from_numpy(yh)
    for make in (lambda: Array.zeros((30, 44), np.float32), lambda: Array.zeros((30, 44), np.float32, order="F"),
                 lambda: Array.zeros((60, 88), np.float32).slice(s_[::2, ::-2])):
        into = make()
        Zip.from_(x).and_(y.t()).map_assign_into(into, Op.MUL)
        want = np.zeros((30, 44), np.float32)
        assert oracle.zip_(4, want, [xh, yh.T]) == 0
        assert_bit_equal(into.to_numpy(), want, "map_assign_into")
    with pytest.raises(_ffi.ShapeError, match="Zip: Producer dimension mismatch"):
        Zip.from_(x).and_(y.t()).map_assign_into(Array.zeros((30, 45), np.float32), Op.MUL)
    # par_map_assign_into is the same call (src/parallel/impl_par_methods.rs:126)
    into = Array.zeros((30, 44), np.float32)
    Zip.from_(x).and_(y.t()).par_map_assign_into(into, Op.ADD)
    assert_bit_equal(into.to_numpy(), xh + yh.T)


def test_scan_keeps_the_first_element(nd):
    """accumulate_axis_inplace leaves element 0 untouched (src/impl_methods.rs:3243 starts at index 1): a leading -0.0
    must stay -0.0 on every scan path (no identity folded in), and a zero stride on the scanned axis is refused."""
    from ndarray_b200 import Array, Op
    from ndarray_b200 import _ffi
    for shape, axis in (((6, 40), 1), ((6, 4096), 1), ((40, 6), 0), ((4096, 16), 0), ((5, 130), 1)):
        h = np.full(shape, -0.0, np.float32)
        d = Array.from_numpy(h)
        d.accumulate_axis_inplace(axis, Op.ADD)
        got = d.to_numpy()
        assert np.all(np.signbit(got)), (shape, axis)  # -0.0 + -0.0 = -0.0 all the way
        h2 = np.ones(shape, np.float64)
        idx = [slice(None)] * 2
        idx[axis] = 0
        h2[tuple(idx)] = -0.0
        d2 = Array.from_numpy(h2)
        d2.accumulate_axis_inplace(axis, Op.ADD)
        g2 = d2.to_numpy()
        assert np.all(np.signbit(g2[tuple(idx)])) and np.array_equal(g2, np.cumsum(h2, axis=axis))
    b = Array.from_numpy(np.arange(8, dtype=np.float32)).broadcast((5, 8))
    with pytest.raises(_ffi.NdbError):
        b.accumulate_axis_inplace(0, Op.ADD)


# ---- BASELINE sizes ------------------------------------------------------------------------------------------------
def test_config1_full_size_against_the_oracle(nd, oracle):
    """C1 at 4096^2: a = b*c + a bit-exact against the oracle's Zip, sum_axis on both axes within 1e-5*sqrt(n)."""
    from ndarray_b200 import Array, Axis, Zip
    n = 4096
    a, b, c = (oracle.counter_fill(np.empty((n, n), np.float32), seed, 0, 8) for seed in (0x5EED0001, 0x5EED0002, 0x5EED0003))
    da, db, dc = Array.from_numpy(a), Array.from_numpy(b), Array.from_numpy(c)
    Zip.from_(da).and_(db).and_(dc).for_each("a=b*c+a")
    want = a.copy()
    assert oracle.zip_(7, want, [b, c, want], threads=8) == 0
    assert_bit_equal(da.to_numpy(), want, "C1 a=b*c+a at 4096^2")
    for ax in (0, 1):
        st, ref = oracle.reduce_axis(0, want, ax)  # ndarray's own f32 association
        exact = want.astype(np.float64).sum(axis=ax)
        got = da.sum_axis(Axis(ax)).to_numpy()
        tol = 1e-5 * np.sqrt(n) * np.abs(want).mean()
        assert np.max(np.abs(got - exact)) <= tol, ("device vs f64", ax)
        assert np.max(np.abs(got - ref)) <= 2 * tol, ("device vs ndarray-order oracle", ax)
    tot = float(da.sum())
    assert abs(tot - want.astype(np.float64).sum()) <= 1e-5 * n * np.abs(want).mean()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_config4_full_size_gemm(nd, oracle, dtype, tol):
    """C4 at 8192^3 with (alpha, beta) = (-2.3, pi), CCC and FFC operand orders (tests/oper.rs:595-635 constants):
    the full matrix against OpenBLAS (numpy), and 64 sampled rows against the oracle's Kahan reference
    (crates/numeric-tests/tests/accuracy.rs:143-165).  Tolerance: max-abs-diff / max-abs-ref <= tol (1e-12 x sqrt(k)-scaled
    for f64, 1e-5 for f32 - this pins the 3xTF32 chunked-TMEM accumulation at K = 8192)."""
    from ndarray_b200 import Array
    n = 8192
    alpha, beta = -2.3, np.pi
    A = oracle.counter_fill(np.empty((n, n), dtype), 41, 0, 8)
    B = oracle.counter_fill(np.empty((n, n), dtype), 42, 0, 8)
    C0 = oracle.counter_fill(np.empty((n, n), dtype), 43, 0, 8)
    full = (alpha * (A @ B).astype(np.float64) + beta * C0.astype(np.float64))  # OpenBLAS, same dtype accumulate
    scale = np.abs(full).max()
    rows = np.random.default_rng(1).choice(n, 64, replace=False)
    kahan = np.ascontiguousarray(C0[rows]).copy()
    assert oracle.gemm(alpha, np.ascontiguousarray(A[rows]), B, beta, kahan, mode=1, threads=8) == 0
    # OpenBLAS itself against Kahan on those rows bounds what "agreement with OpenBLAS" can mean
    blas_err = np.abs(full[rows] - kahan).max() / scale
    lim = tol * (np.sqrt(n) if dtype == np.float64 else 1.0)
    for orders in ("CCC", "FFC"):
        dA = Array.from_numpy(np.asfortranarray(A) if orders[0] == "F" else A)
        dB = Array.from_numpy(np.asfortranarray(B) if orders[1] == "F" else B)
        dC = Array.from_numpy(C0)
        nd.general_mat_mul(alpha, dA, dB, beta, dC)
        got = dC.to_numpy().astype(np.float64)
        err_rows = np.abs(got[rows] - kahan).max() / scale
        err_full = np.abs(got - full).max() / scale
        assert err_rows <= lim, (orders, "vs Kahan rows", err_rows)
        assert err_full <= lim + blas_err, (orders, "vs OpenBLAS full", err_full, blas_err)


def test_zip_fold_while(nd):
    """Zip::fold_while (src/zip/mod.rs:640-655): fold in traversal order until the first value that satisfies the stop
    predicate; FoldWhile::Done carries the accumulator of the prefix, FoldWhile::Continue the whole fold.  The order is
    C for standard-layout producers and F when they prefer it (zip/mod.rs:278-281)."""
    from ndarray_b200 import Array, Zip, Op
    rng = np.random.default_rng(17)
    for dtype in (np.int64, np.float64, np.float32, np.int32):
        h = rng.integers(1, 50, (37, 53)).astype(dtype)
        h[20, 11] = -5
        h[30, 2] = -7
        for order in ("C", "F"):
            hh = np.asfortranarray(h) if order == "F" else h
            d = Array.from_numpy(hh)
            seq = hh.ravel(order=order)  # Zip's traversal order for this layout
            first = int(np.argmax(seq < 0))
            acc, done = Zip.from_(d).fold_while(dtype(3), Op.COPY, "sum", 27, 0)  # 27 = LT: stop at the first x < 0
            assert done and acc == dtype(3) + seq[:first].sum(dtype=dtype), (dtype, order)
            acc, done = Zip.from_(d).fold_while(None, Op.COPY, "max", 27, 0)
            assert done and acc == seq[:first].max()
            acc, done = Zip.from_(d).fold_while(dtype(0), Op.COPY, "sum", 27, -100)  # never fires: Continue(whole fold)
            assert not done and acc == seq.sum(dtype=dtype)
            # two producers: stop where a*b < 0
            acc, done = Zip.from_(d).and_(d.view()).fold_while(dtype(0), Op.ADD, "sum", 27, 0)
            assert done and acc == (2 * seq[:first]).sum(dtype=dtype)
        first_elem = Array.from_numpy(np.array([-1, 2, 3], dtype=dtype))
        acc, done = Zip.from_(first_elem).fold_while(dtype(9), Op.COPY, "sum", 27, 0)
        assert done and acc == 9  # Done before anything was folded: the initial accumulator
    e = Array.zeros((0, 3), np.float32)
    acc, done = Zip.from_(e).fold_while(np.float32(1.5), Op.COPY, "sum", 27, 0)
    assert not done and acc == 1.5


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32])
def test_chained_single_pass_scan(nd, dtype):
    """The chained scan (long strided axis, few dense lanes: 64-row chunks passed from block to block through epoch-tagged
    packets) against numpy: integers exact, floats to rounding; repeated launches reuse the packets (epochs), an outer kept
    axis rides along, and a backwards scanned axis is walked in logical order."""
    from ndarray_b200 import Array, Op, s_
    rng = np.random.default_rng(41)
    kind = np.dtype(dtype).kind
    w = 16 // np.dtype(dtype).itemsize
    for shape, axis in (((1024, 32 * w * 3), 0), ((3, 640, 32 * w), 1), ((4096, 32 * w), 0)):
        h = (rng.integers(-3, 4, shape) if kind == "i" else 1.0 + 0.001 * rng.standard_normal(shape)).astype(dtype)
        for op, ref in ((Op.ADD, np.cumsum), (Op.MUL, np.cumprod)):
            for rep in range(3):
                d = Array.from_numpy(h)
                d.accumulate_axis_inplace(axis, op)
                got = d.to_numpy()
                want = ref(h.astype(np.float64) if kind == "f" else h, axis=axis)
                if kind == "i":
                    assert np.array_equal(got, want.astype(dtype)), (shape, axis, rep)
                else:
                    tol = (1e-12 if dtype == np.float64 else 1e-5) * np.sqrt(shape[axis]) * max(1.0, float(np.abs(want).max()))
                    assert np.max(np.abs(got - want)) <= tol, (shape, axis, rep)
    # reversed scanned axis and a leading -0.0 column
    h = np.ones((1024, 32 * w), dtype)
    d = Array.from_numpy(h)
    d.slice(s_[::-1, :]).accumulate_axis_inplace(0, Op.ADD)
    assert np.array_equal(d.to_numpy(), np.cumsum(h[::-1], axis=0)[::-1])
    if kind == "f":
        z = np.full((1024, 32 * w), -0.0, dtype)
        dz = Array.from_numpy(z)
        dz.accumulate_axis_inplace(0, Op.ADD)
        assert np.all(np.signbit(dz.to_numpy()))


def test_swap_axes_insert_axis_and_azip(nd, oracle):
    """The axis transforms that had no test in round 1 (src/impl_methods.rs:2487 swap_axes, :2700 insert_axis) and the
    azip! / par_azip! sugar (src/zip/zipmacro.rs:98): views must address exactly the elements numpy's do, and Zips over
    them must match the oracle bit for bit."""
    from ndarray_b200 import Array, NewAxis, Zip, azip, par_azip, s_
    rng = np.random.default_rng(71)
    h = rng.integers(-1000, 1000, (5, 7, 9)).astype(np.int64)
    d = Array.from_numpy(h)
    for a, b in ((0, 2), (1, 2), (0, 0)):
        v = d.swap_axes(a, b)
        assert v.shape == np.swapaxes(h, a, b).shape and v.strides == tuple(s // 8 for s in np.swapaxes(h, a, b).strides)
        assert np.array_equal(v.to_numpy(), np.swapaxes(h, a, b))
    for ax in (0, 1, 3):
        v = d.insert_axis(ax)
        assert v.shape == np.expand_dims(h, ax).shape
        assert np.array_equal(v.to_numpy(), np.expand_dims(h, ax))
        # the new unit axis broadcasts like any other: out[.., j, ..] = v for j in 0..3
        shape = list(v.shape)
        shape[ax] = 4
        out = Array.zeros(tuple(shape), np.int64)
        Zip.from_(out).and_broadcast(v).for_each("a=b")
        assert np.array_equal(out.to_numpy(), np.broadcast_to(np.expand_dims(h, ax), shape))
    # s![.., NewAxis, ..] goes through insert_axis as well
    v = d.slice(s_[:, NewAxis, ::2, :])
    assert v.shape == h[:, np.newaxis, ::2, :].shape and np.array_equal(v.to_numpy(), h[:, np.newaxis, ::2, :])
    # azip!((a in &mut a, &b in &b, &c in &c) *a = b * c + *a) on swapped views
    fa, fb, fc = (rng.standard_normal((6, 8)).astype(np.float32) for _ in range(3))
    da, db, dc = Array.from_numpy(fa), Array.from_numpy(np.ascontiguousarray(fb.T)), Array.from_numpy(fc)
    azip("a=b*c+a", da, db.swap_axes(0, 1), dc)
    want = fa.copy()
    assert oracle.zip_(7, want, [fb, fc, want]) == 0
    assert_bit_equal(da.to_numpy(), want, "azip on a swapped view")
    par_azip("a+=b", da, dc)
    assert oracle.zip_(2, want, [want, fc]) == 0
    assert_bit_equal(da.to_numpy(), want, "par_azip")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_zip_for_each_fold_fused(nd, oracle, dtype):
    """Zip::fold whose closure writes its first producer (ndb_zip_fold_write): the written array is bit-exact against the
    oracle's Zip and the folded value equals the fold of that array - dense producers (the fused kernel, ragged tail
    included), strided ones (Zip + reduction), sum / min / max."""
    from ndarray_b200 import Array, Zip, Op, s_
    rng = np.random.default_rng(99)
    kind = np.dtype(dtype).kind
    for shape in ((1000003,), (257, 129), (64, 64)):
        def mk():
            return (rng.integers(-1000, 1000, shape) if kind == "i" else rng.uniform(-1, 1, shape)).astype(dtype)
        ah, bh, ch = mk(), mk(), mk()
        for views in ("dense", "strided"):
            if views == "strided" and len(shape) == 1:
                continue
            sl = (lambda x: x) if views == "dense" else (lambda x: x.slice(s_[::-1, ::2]))
            hs = (lambda x: x) if views == "dense" else (lambda x: x[::-1, ::2])
            a, b, c = Array.from_numpy(ah), Array.from_numpy(bh), Array.from_numpy(ch)
            got = Zip.from_(sl(a)).and_(sl(b)).and_(sl(c)).for_each_fold("a=b*c+a", "sum")
            want = ah.copy()
            wv = hs(want)
            assert oracle.zip_(7, wv, [hs(bh), hs(ch), wv]) == 0
            assert_bit_equal(a.to_numpy(), want, f"for_each_fold {views}: written array")
            if kind == "i":
                with np.errstate(over="ignore"):
                    assert got == wv.sum(dtype=dtype)
            else:
                tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(wv.size) * max(np.abs(wv).mean(), 1e-30)
                assert abs(float(got) - wv.astype(np.float64).sum()) <= tol
            for fold, fn in (("min", np.min), ("max", np.max)):
                a2 = Array.from_numpy(ah)
                g2 = Zip.from_(sl(a2)).and_(sl(b)).and_(sl(c)).for_each_fold("a=b+c", fold)
                assert g2 == fn(hs(bh) + hs(ch))
                assert_bit_equal(sl(a2).to_numpy(), (hs(bh) + hs(ch)).astype(dtype), "for_each_fold min/max: written array")


def test_config2_full_size_against_the_oracle(nd, oracle):
    """C2 at 512^3, every view case of SURVEY 8d, WHOLE outputs bit-exact against the oracle's Zip traversal (threads only
    split the outer axis, the arithmetic per element is ndarray's): (i) x.t() + y, (ii) s![..;-2,..,..] + s![..;2,..,..] and the
    inner-axis variant s![..,..,..;-2], (iii) x * row.broadcast, (iv) all three views in one Zip."""
    from ndarray_b200 import Array, Zip, s_
    n = 512
    xh = oracle.counter_fill(np.empty((n, n, n), np.float32), 11, 0, 8)
    yh = oracle.counter_fill(np.empty((n, n, n), np.float32), 12, 0, 8)
    rowh = oracle.counter_fill(np.empty(n, np.float32), 13, 0, 1)
    x, y, row = Array.from_numpy(xh), Array.from_numpy(yh), Array.from_numpy(rowh)
    out = Array.uninit((n, n, n), np.float32)
    want = np.empty((n, n, n), np.float32)
    Zip.from_(out).and_(x.t()).and_(y).for_each("a=b+c")
    assert oracle.zip_(2, want, [xh.transpose(2, 1, 0), yh], threads=8) == 0
    assert_bit_equal(out.to_numpy(), want, "C2 (i) transposed")
    half, wh = out.slice(s_[: n // 2, :, :]), want[: n // 2]
    Zip.from_(half).and_(x.slice(s_[::-2, :, :])).and_(y.slice(s_[::2, :, :])).for_each("a=b+c")
    assert oracle.zip_(2, wh, [xh[::-1][::2], yh[::2]], threads=8) == 0
    assert_bit_equal(half.to_numpy(), wh, "C2 (ii) negative outer step")
    inner, wi = out.slice(s_[:, :, : n // 2]), want[:, :, : n // 2]
    Zip.from_(inner).and_(x.slice(s_[:, :, ::-2])).and_(y.slice(s_[:, :, ::2])).for_each("a=b+c")
    assert oracle.zip_(2, wi, [xh[:, :, ::-1][:, :, ::2], yh[:, :, ::2]], threads=8) == 0
    assert_bit_equal(inner.to_numpy(), wi, "C2 (ii) negative inner step")
    Zip.from_(out).and_(x).and_broadcast(row).for_each("a=b*c")
    assert oracle.zip_(4, want, [xh, np.broadcast_to(rowh, (n, n, n))], threads=8) == 0
    assert_bit_equal(out.to_numpy(), want, "C2 (iii) broadcast row")
    Zip.from_(half).and_(y.slice(s_[::2, :, :])).and_(row.broadcast((n // 2, n, n))).and_(x.t().slice(s_[::-2, :, :])).for_each("a=b*c+d")
    assert oracle.zip_(7, wh, [yh[::2], np.broadcast_to(rowh, (n // 2, n, n)), xh.transpose(2, 1, 0)[::-1][::2]], threads=8) == 0
    assert_bit_equal(half.to_numpy(), wh, "C2 (iv) all three views in one Zip")


def test_config3_full_size_against_the_oracle(nd, oracle):
    """C3 at 16384^2 against the ORACLE (ndarray's own algorithms, not numpy): i64 sum_axis and fold_axis(min) on both axes
    bit-exact; f64 sum_axis / mean_axis / var_axis(ddof=1) on both axes within 1e-12 * sqrt(n) * scale of the oracle's
    reference-order result (Axis(1) of the oracle's Welford strides through memory: a few seconds)."""
    from ndarray_b200 import Array, Axis
    n = 16384
    hi = oracle.counter_fill(np.empty((n, n), np.int64), 21, 0, 8)
    a = Array.from_numpy(hi)
    for ax in (0, 1):
        st, want = oracle.reduce_axis(0, hi, ax)
        assert st == 0 and np.array_equal(a.sum_axis(Axis(ax)).to_numpy(), want), ("i64 sum_axis", ax)
        assert np.array_equal(a.fold_axis(Axis(ax), np.iinfo(np.int64).max, "min").to_numpy(), hi.min(axis=ax)), ("i64 fold min", ax)
    del a, hi
    hf = oracle.counter_fill(np.empty((n, n), np.float64), 22, 0, 8)
    f = Array.from_numpy(hf)
    tol = 1e-12 * np.sqrt(n)
    for ax in (0, 1):
        st, want = oracle.reduce_axis(0, hf, ax)
        assert st == 0 and np.max(np.abs(f.sum_axis(Axis(ax)).to_numpy() - want)) <= tol * 8 * np.abs(hf).mean() * np.sqrt(n), ("f64 sum_axis", ax)
        st, want = oracle.reduce_axis(4, hf, ax)
        assert st == 0 and np.max(np.abs(f.mean_axis(Axis(ax)).to_numpy() - want)) <= tol, ("f64 mean_axis", ax)
    for ax in (0, 1):
        st, want = oracle.reduce_axis(5, hf, ax, 1.0)
        assert st == 0 and np.max(np.abs(f.var_axis(Axis(ax), 1.0).to_numpy() - want)) <= tol, ("f64 var_axis", ax)


def test_counter_generator_against_the_committed_fixture_on_device(nd):
    """The device generator against tests/golden/counter_rng.json (independent pure-Python splitmix64), all element types."""
    import json
    from pathlib import Path
    from ndarray_b200 import Array
    from ndarray_b200._ffi import lib, check
    fx = json.loads((Path(__file__).parent / "golden" / "counter_rng.json").read_text())
    ctx = nd.default_context()
    for c in fx["cases"]:
        n = len(c["i64"])
        for dt, key in ((np.float32, "f32_bits"), (np.float64, "f64"), (np.int32, "i32"), (np.int64, "i64")):
            d = Array.uninit((n,), dt, ctx=ctx)
            check(lib().ndb_fill_counter_rng(ctx.handle, C.byref(d.v), c["seed"], c["first"]))
            got = d.to_numpy()
            if key == "f32_bits":
                assert got.view(np.uint32).tolist() == c[key], c
            elif key == "f64":
                assert [float(x).hex() for x in got] == c[key], c
            else:
                assert got.tolist() == c[key], c


def test_scan_inside_a_captured_graph_replays_on_new_data(nd):
    """A captured running sum must give the right answer on every replay: the chained kernel's epoch is a launch argument
    (a replay would find last time's packets valid), so a capturing stream takes the two-pass form."""
    from ndarray_b200 import Array, Op
    ctx = nd.default_context()
    rng = np.random.default_rng(43)
    shape = (4096, 64)
    src = Array.from_numpy(rng.integers(-5, 6, shape).astype(np.int64))
    d = Array.zeros(shape, np.int64)
    d.assign(src)
    d.accumulate_axis_inplace(0, Op.ADD)  # outside the capture: scratch allocations
    with ctx.capture() as g:
        d.assign(src)
        d.accumulate_axis_inplace(0, Op.ADD)
    for rep in range(3):
        h = rng.integers(-5, 6, shape).astype(np.int64)
        src.assign(Array.from_numpy(h))
        g.launch()
        ctx.sync()
        assert np.array_equal(d.to_numpy(), np.cumsum(h, axis=0)), rep
    g.close()
