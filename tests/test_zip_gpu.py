"""GPU parity: the Zip engine (ndb_zip through the C ABI) against the oracle, bit-exact.

Views are built once in numpy and mirrored on the device with identical (offset, shape, strides), so
both sides traverse the same elements of the same buffers."""
import zlib

import numpy as np
import pytest

from .conftest import assert_bit_equal, device_view_of

pytestmark = pytest.mark.gpu

OP = dict(COPY=0, FILL=1, ADD=2, SUB=3, MUL=4, DIV=5, REM=6, MUL_ADD=7, ADD_S=8, SUB_S=9, MUL_S=10, DIV_S=11, REM_S=12,
          RSUB_S=13, RDIV_S=14, NEG=15, AXPY=16, ADD_N=17, ABS=18, SQRT=19, MIN=20, MAX=21, SCALE_ADD=22)
ARITY = dict(COPY=1, FILL=0, ADD=2, SUB=2, MUL=2, DIV=2, REM=2, MUL_ADD=3, ADD_S=1, SUB_S=1, MUL_S=1, DIV_S=1, REM_S=1,
             RSUB_S=1, RDIV_S=1, NEG=1, AXPY=2, ABS=1, SQRT=1, MIN=2, MAX=2, SCALE_ADD=2)
NSCAL = dict(FILL=1, ADD_S=1, SUB_S=1, MUL_S=1, DIV_S=1, REM_S=1, RSUB_S=1, RDIV_S=1, AXPY=1, SCALE_ADD=2)
DTYPES = [np.float32, np.float64, np.int32, np.int64]


def rand(rng, shape, dtype, nonzero=False):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        a = rng.uniform(-1, 1, size=shape).astype(dtype)
        if nonzero:
            a = np.where(np.abs(a) < 0.05, dtype.type(0.5), a)
    else:
        info = np.iinfo(dtype)
        a = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
        if nonzero:
            a = np.where((a == 0) | (a == -1), dtype.type(3), a)
    return a


GUARD = 64  # guard elements either side of every mirrored buffer (compute-sanitizer is not available on the GPU pool)


class Mirror:
    """A host buffer and its device copy; views are taken on both with the same geometry.  The device copy sits between
    two guard bands holding a sentinel; `download` fails if a kernel wrote outside the buffer."""

    def __init__(self, nd, host: np.ndarray):
        self.nd = nd
        self.host = np.array(host, copy=True, order="C")
        self.upload()

    def upload(self):
        flat = self.host.reshape(-1)
        padded = np.empty(flat.size + 2 * GUARD, dtype=self.host.dtype)
        padded.view(np.uint8)[...] = 0x5A
        padded[GUARD:GUARD + flat.size] = flat
        self.padded_dev = self.nd.Array.from_numpy(padded)
        # the device array the tests see: same shape/strides as the host buffer, inside the guards
        self.dev = self.padded_dev.raw_view_with(GUARD, self.host.shape, [st // self.host.itemsize for st in self.host.strides])

    def view(self, fn):
        hv = fn(self.host)
        item = self.host.itemsize
        off = (hv.__array_interface__["data"][0] - self.host.__array_interface__["data"][0]) // item
        return hv, self.padded_dev.raw_view_with(GUARD + off, hv.shape, [st // item for st in hv.strides])

    def download(self):
        padded = self.padded_dev.to_numpy()
        g = padded.view(np.uint8)
        gb = GUARD * self.host.itemsize
        assert (g[:gb] == 0x5A).all() and (g[g.size - gb:] == 0x5A).all(), "kernel wrote outside the buffer (guard band hit)"
        return padded[GUARD:GUARD + self.host.size].reshape(self.host.shape)


def run_case(nd, oracle, opname, dtype, shape, out_fn, in_fns, rng, scalars=None):
    """out_fn / in_fns map a freshly allocated base array of `base_shape` to the view used."""
    op = OP[opname]
    nonzero = opname in ("DIV", "REM", "RDIV_S")
    mo = Mirror(nd, rand(rng, shape["out"], dtype))
    mins = [Mirror(nd, rand(rng, shp, dtype, nonzero=nonzero and i == len(in_fns) - 1 or opname == "RDIV_S"))
            for i, shp in enumerate(shape["ins"])]
    if opname == "SQRT":
        for m in mins:
            m.host[...] = np.abs(m.host)
            m.upload()
    ho, do = mo.view(out_fn)
    hins, dins = [], []
    for m, fn in zip(mins, in_fns):
        h, d = m.view(fn)
        hins.append(h)
        dins.append(d)
    if scalars is None:
        scalars = []
        for _ in range(NSCAL.get(opname, 0)):
            scalars.append(rand(rng, (), dtype, nonzero=True)[()])
    st = oracle.zip_(op, ho, hins, scalars)
    assert st == 0, st
    from ndarray_b200.array import _zip_call
    _zip_call(do.ctx, op, do, dins, scalars)
    assert_bit_equal(mo.download(), mo.host, f"{opname} {np.dtype(dtype).name} {shape}")


ident = lambda a: a


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("opname", [k for k in ARITY if k != "ADD_N"])
def test_every_op_contiguous(nd, oracle, opname, dtype):
    if opname == "SQRT" and np.dtype(dtype).kind != "f":
        pytest.skip("sqrt is float-only")
    rng = np.random.default_rng(zlib.crc32(f"{opname}{np.dtype(dtype).name}".encode()))
    n = 1000 * 37 + 3  # not a multiple of the vector width
    run_case(nd, oracle, opname, dtype, dict(out=(n,), ins=[(n,)] * ARITY[opname]), ident, [ident] * ARITY[opname], rng)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("opname", ["ADD", "MUL_ADD", "DIV", "AXPY"])
def test_ops_on_strided_views(nd, oracle, opname, dtype):
    rng = np.random.default_rng(11)
    k = ARITY[opname]
    shp = (37, 50, 66)
    # every producer a different kind of view of a (37, 50, 66)-shaped logical array
    fns = [lambda a: a[::-1], lambda a: a.transpose(2, 1, 0), lambda a: a[:, ::2, ::-1][:, :, ::1]]
    base_shapes = [shp, shp[::-1], (37, 100, 66)]
    run_case(nd, oracle, opname, dtype, dict(out=shp, ins=base_shapes[:k]), ident, fns[:k], rng)


VIEW_CASES = {
    # name: (out base shape, out view fn, [(in base shape, in view fn), ...]) for a binary ADD
    "c_c_c": ((64, 96), ident, [((64, 96), ident), ((64, 96), ident)]),
    "f_order_all": ((96, 64), lambda a: a.T, [((96, 64), lambda a: a.T), ((96, 64), lambda a: a.T)]),
    "transposed_in_2d": ((130, 70), ident, [((70, 130), lambda a: a.T), ((130, 70), ident)]),
    "transposed_in_3d": ((70, 33, 65), ident, [((65, 33, 70), lambda a: a.transpose(2, 1, 0)), ((70, 33, 65), ident)]),
    "two_transposed": ((70, 65), ident, [((65, 70), lambda a: a.T), ((65, 70), lambda a: a.T)]),
    "neg_outer": ((40, 50), ident, [((80, 50), lambda a: a[::-2]), ((80, 50), lambda a: a[::2])]),
    "neg_inner": ((40, 64), ident, [((40, 128), lambda a: a[:, ::-2]), ((40, 64), lambda a: a[:, ::-1])]),
    "rev_inner_vec": ((40, 64), ident, [((40, 64), lambda a: a[:, ::-1]), ((40, 64), ident)]),
    "rev_inner_3d": ((9, 20, 128), ident, [((9, 20, 128), lambda a: a[::-1, :, ::-1]), ((9, 20, 128), lambda a: a[:, :, ::-1])]),
    "rev_inner_unaligned": ((40, 63), ident, [((40, 63), lambda a: a[:, ::-1]), ((40, 63), ident)]),
    "out_reversed": ((40, 64), lambda a: a[::-1, ::-1], [((40, 64), ident), ((40, 64), lambda a: a[::-1])]),
    "out_strided": ((40, 128), lambda a: a[:, ::2], [((40, 64), ident), ((40, 64), ident)]),
    "out_f_in_c": ((64, 40), lambda a: a.T, [((40, 64), ident), ((40, 64), ident)]),
    "bcast_row": ((33, 70, 64), ident, [((33, 70, 64), ident), ((64,), lambda a: np.broadcast_to(a, (33, 70, 64)))]),
    "bcast_col": ((33, 64), ident, [((33, 1), lambda a: np.broadcast_to(a, (33, 64))), ((33, 64), ident)]),
    "bcast_scalar": ((33, 64), ident, [((), lambda a: np.broadcast_to(a, (33, 64))), ((33, 64), ident)]),
    "bcast_mid": ((8, 16, 32), ident, [((8, 1, 32), lambda a: np.broadcast_to(a, (8, 16, 32))), ((8, 16, 32), ident)]),
    "offset_unaligned": ((1001,), lambda a: a[1:], [((1001,), lambda a: a[:-1]), ((1003,), lambda a: a[3:])]),
    "offset_rows": ((17, 67), lambda a: a[1:, 3:], [((17, 67), lambda a: a[:-1, :-3]), ((16, 64), ident)]),
    "zero_d": ((), ident, [((), ident), ((), ident)]),
    "one_elem": ((1, 1, 1), ident, [((1, 1, 1), ident), ((1,), lambda a: np.broadcast_to(a, (1, 1, 1)))]),
    "empty": ((3, 0, 5), ident, [((3, 0, 5), ident), ((3, 0, 5), ident)]),
    "five_d": ((3, 4, 5, 6, 7), ident, [((7, 6, 5, 4, 3), lambda a: a.transpose(4, 3, 2, 1, 0)), ((3, 4, 5, 6, 14), lambda a: a[..., ::2])]),
    "short_inner": ((5000, 3), ident, [((3, 5000), lambda a: a.T), ((5000, 3), ident)]),
    "long_rows": ((3, 40000), ident, [((3, 40000), lambda a: a[::-1]), ((3, 80000), lambda a: a[:, ::2])]),
    "tile_ragged": ((67, 129), ident, [((129, 67), lambda a: a.T), ((67, 129), ident)]),
    "tile_step2": ((70, 80), ident, [((80, 140), lambda a: a.T[::2]), ((70, 80), ident)]),
    "tile_step_neg2": ((70, 80), ident, [((80, 140), lambda a: a.T[::-2]), ((70, 80), ident)]),
    "tile_rev_vec": ((64, 128), ident, [((128, 64), lambda a: a.T[::-1]), ((64, 128), ident)]),
    "tile_3d_step": ((32, 20, 64), ident, [((64, 20, 64), lambda a: a.transpose(2, 1, 0)[::-2]), ((32, 20, 64), ident)]),
    "tile_vec_step2": ((64, 128), ident, [((128, 128), lambda a: a.T[::2]), ((64, 128), ident)]),
    "tile_vec_step_neg2": ((64, 128), ident, [((128, 128), lambda a: a.T[::-2]), ((64, 128), ident)]),
    "tile_vec_two_staged": ((64, 128), ident, [((128, 128), lambda a: a.T[::-2]), ((128, 64), lambda a: a.T)]),
    "tile_reversed_dense_axis": ((70, 80), ident, [((80, 70), lambda a: a.T[::-1]), ((70, 80), ident)]),
}


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("case", sorted(VIEW_CASES))
def test_view_geometries(nd, oracle, case, dtype):
    oshape, ofn, ins = VIEW_CASES[case]
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    run_case(nd, oracle, "ADD", dtype, dict(out=oshape, ins=[s for s, _ in ins]), ofn, [f for _, f in ins], rng)


def test_kernel_classes_are_exercised(nd):
    """The geometry table above really reaches every kernel class (guards against a planner that
    silently sends everything to the generic kernel)."""
    seen = set()
    for case, (oshape, ofn, ins) in VIEW_CASES.items():
        mo = Mirror(nd, np.zeros(oshape, np.float32))
        _, do = mo.view(ofn)
        dins = []
        for shp, fn in ins:
            m = Mirror(nd, np.zeros(shp, np.float32))
            dins.append(m.view(fn)[1])
        z = nd.Zip.from_(do)
        for d in dins:
            z = z.and_(d)
        seen.add(nd.KERNEL_NAMES[z.plan(nd.Op.ADD).kernel])
    assert seen >= {"noop", "flat_vec", "flat", "rows_vec", "rows", "tile"}, seen


@pytest.mark.parametrize("n_in", [2, 3, 5, 8])
def test_add_n(nd, oracle, n_in):
    rng = np.random.default_rng(n_in)
    shp = (33, 65)
    run_case(nd, oracle, "ADD_N", np.int32, dict(out=shp, ins=[shp] * n_in), ident, [ident] * n_in, rng)
    run_case(nd, oracle, "ADD_N", np.float32, dict(out=shp, ins=[shp[::-1]] * n_in), ident, [lambda a: a.T] * n_in, rng)


def test_in_place_config1_form(nd, oracle):
    """Zip::from(&mut a).and(&b).and(&c).for_each(|a,&b,&c| *a = b*c + *a): `a` is read and written."""
    rng = np.random.default_rng(5)
    n = 1 << 20
    a, b, c = (rand(rng, (1024, 1024), np.float32) for _ in range(3))
    want = a.copy()
    assert oracle.zip_(OP["MUL_ADD"], want, [b, c, want]) == 0
    assert np.array_equal(want, (b * c).astype(np.float32) + a)  # two roundings, no fma
    da, db, dc = (nd.Array.from_numpy(x) for x in (a, b, c))
    nd.Zip.from_(da).and_(db).and_(dc).for_each("a=b*c+a")
    assert_bit_equal(da.to_numpy(), want, "a=b*c+a")
    assert n == da.len()


def test_special_float_values(nd, oracle):
    vals = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, -3.4e38, 1.1754944e-38,
                     7.0, -7.5, 0.1, 2.0 ** -126], np.float32)
    x, y = np.meshgrid(vals, vals)
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    for opname in ["ADD", "SUB", "MUL", "DIV", "REM", "MIN", "MAX"]:
        want = np.zeros_like(x)
        with np.errstate(all="ignore"):
            assert oracle.zip_(OP[opname], want, [x, y]) == 0
        out = nd.Array.zeros(x.shape, np.float32)
        from ndarray_b200.array import _zip_call
        _zip_call(out.ctx, OP[opname], out, [nd.Array.from_numpy(x), nd.Array.from_numpy(y)])
        assert_bit_equal(out.to_numpy(), want, opname)


def test_integer_division_traps(nd):
    a = nd.Array.from_numpy(np.array([1, -(2 ** 31), 7], np.int32))
    z = nd.Array.from_numpy(np.array([1, 1, 0], np.int32))
    m1 = nd.Array.from_numpy(np.array([1, -1, 1], np.int32))
    with pytest.raises(nd.NdbError) as e:
        _ = a / z
    assert e.value.status == 9
    with pytest.raises(nd.NdbError):
        _ = a / m1
    ok = a / nd.Array.from_numpy(np.array([1, 2, -2], np.int32))  # the flag was reset
    assert ok.to_numpy().tolist() == [1, -(2 ** 30), -3]


def test_dimension_mismatch_panics(nd):  # tests/azip.rs:258-267
    a = nd.Array.zeros((5, 7))
    b = nd.Array.zeros((6, 7))
    with pytest.raises(nd.ShapeError, match="Zip: Producer dimension mismatch"):
        nd.Zip.from_(a).and_(b)
    from ndarray_b200.array import _zip_call
    with pytest.raises(nd.ShapeError, match="Zip: Producer dimension mismatch, expected: \\[5, 7\\], got: \\[6, 7\\]"):
        _zip_call(a.ctx, nd.Op.COPY, a, [b])


def test_reference_golden_vectors_on_device(nd):
    from ndarray_b200 import s_, Array, Zip, Op
    # tests/array.rs:1759-1794
    buf = Array.from_numpy(np.arange(1, 13, dtype=np.int64))
    s = buf.raw_view_with(0, (2, 3, 2), (1, 4, 2))
    a = s.slice(s_[::-1, ::2, :])
    b = s.slice(s_[:2, -1, :])
    assert (a.to_owned() + b).to_numpy().tolist() == [[[11, 15], [20, 24]], [[10, 14], [19, 23]]]
    a2 = s.slice(s_[::-1, ::2, :])
    b2 = s.slice(s_[:, ::2, :])
    assert (a2 + b2).to_numpy().tolist() == [[[3, 7], [19, 23]], [[3, 7], [19, 23]]]
    # src/slice.rs:800-806
    arr = Array.from_numpy(np.array([0, 1, 2, 3], np.int64))
    assert arr.slice(s_[1:3:-1]).to_numpy().tolist() == [2, 1]
    assert arr.slice(s_[1::-2]).to_numpy().tolist() == [3, 1]
    assert arr.slice(s_[::-2]).to_numpy().tolist() == [3, 1]
    # tests/array.rs:1745-1757
    a = Array.from_numpy(np.array([[2], [3], [4]], np.int64))
    b = Array.from_numpy(np.array([5, 6, 7], np.int64))
    assert (a + b).to_numpy().tolist() == [[7, 8, 9], [8, 9, 10], [9, 10, 11]]
    assert (a - b).to_numpy().tolist() == [[-3, -4, -5], [-2, -3, -4], [-1, -2, -3]]
    assert (a * b).to_numpy().tolist() == [[10, 12, 14], [15, 18, 21], [20, 24, 28]]
    assert (b / a).to_numpy().tolist() == [[2, 3, 3], [1, 2, 2], [1, 1, 1]]
    # tests/azip.rs:481-497 (nine producers)
    j = np.arange(62, dtype=np.int32)
    parts = [Array.zeros(62, np.int32)] + [Array.from_numpy(j * k) for k in (1, 2, 4, 8, 16, 32, 64, 128)]
    z = Zip.from_(parts[0])
    for p in parts[1:]:
        z = z.and_(p)
    z.for_each("a=sum(rest)")
    assert parts[0].to_numpy().tolist() == (j * 255).tolist()
    # tests/azip.rs:234-256 (and_broadcast)
    n = 16
    a = Array.zeros((n, n), np.float32)
    bh = (1.0 / (1.0 + 2 * np.arange(n, dtype=np.float32))).reshape(1, n).astype(np.float32)
    b = Array.from_numpy(bh)
    d = Array.from_elem((1, n), 1.0, np.float32)
    e = Array.from_elem((), 2.0, np.float32)
    Zip.from_(a).and_broadcast(b).and_broadcast(d).and_broadcast(e).for_each(Op.ADD_N)
    np.testing.assert_allclose(a.to_numpy(), np.broadcast_to(bh + 3.0, (n, n)), atol=1e-4)
    # tests/azip.rs:53-78: map_collect keeps the inputs' order
    b = Array.from_numpy(np.fromfunction(lambda i, j: 1. / (i + 2 * j + 1), (5, 10)).astype(np.float32))
    c = Array.from_numpy(np.exp(np.add.outer(np.arange(5.), np.arange(10.))).astype(np.float32))
    r = Zip.from_(b).and_(c).map_collect(Op.ADD)
    assert r.strides == b.strides
    r = Zip.from_(b.t()).and_(c.t()).map_collect(Op.ADD)
    assert r.strides == b.t().strides
    np.testing.assert_allclose(r.to_numpy(), (b.to_numpy() + c.to_numpy()).T, atol=1e-6)
    # tests/azip.rs:472-478: Zip over an empty shape is a no-op
    e0 = Array.zeros(0, np.float32)
    Zip.from_(e0).and_(Array.zeros(0, np.float32)).for_each("a=b")


def test_config2_full_size_properties(nd):
    """BASELINE config 2 at full size (f32 512^3): no full-size oracle pass (too slow for a test), so use
    size-independent properties: x.t().t() + 0 == x, reversal twice is identity, broadcast-row product
    against a row-wise numpy check on sampled planes."""
    from ndarray_b200 import s_, Array, Zip, Op
    n = 512
    rng = np.random.default_rng(2)
    xh = rng.uniform(-1, 1, size=(n, n, n)).astype(np.float32)
    x = Array.from_numpy(xh)
    y = Array.zeros((n, n, n), np.float32)
    out = Array.uninit((n, n, n), np.float32)
    Zip.from_(out).and_(x.t()).and_(y).for_each("a=b+c")          # (i) transposed add
    assert nd.KERNEL_NAMES[Zip.from_(out).and_(x.t()).and_(y).plan(Op.ADD).kernel] == "tile"
    for i in (0, 1, 255, 511):
        assert np.array_equal(out.index_axis(0, i).to_numpy(), xh[:, :, i].T)
    back = Array.uninit((n, n, n), np.float32)
    Zip.from_(back).and_(out.t()).and_(y).for_each("a=b+c")       # transpose twice == identity
    assert float((back - x).mapv(Op.ABS).sum()) == 0.0
    half = Array.uninit((n // 2, n, n), np.float32)               # (ii) negative outer step
    Zip.from_(half).and_(x.slice(s_[::-2, :, :])).and_(y.slice(s_[::2, :, :])).for_each("a=b+c")
    for i in (0, 100, 255):
        assert np.array_equal(half.index_axis(0, i).to_numpy(), xh[n - 1 - 2 * i])
    rowh = rng.uniform(-1, 1, size=n).astype(np.float32)          # (iii) broadcast row
    row = Array.from_numpy(rowh)
    Zip.from_(out).and_(x).and_broadcast(row).for_each("a=b*c")
    for i in (0, 300, 511):
        assert np.array_equal(out.index_axis(0, i).to_numpy(), xh[i] * rowh[None, :])


def test_zip_all_any_fold(nd):  # tests/azip.rs:460-478
    from ndarray_b200 import Array, Zip, Op
    a = Array.zeros(62, np.float32)
    b = Array.ones(62, np.float32)
    ch = np.ones(62, np.float32)
    ch[5] = 0.0
    c = Array.from_numpy(ch)
    one = Array.from_elem((), 1.0, np.float32).broadcast(62)
    s_ab = Zip.from_(a).and_(b).map_collect(Op.ADD)
    s_ac = Zip.from_(a).and_(c).map_collect(Op.ADD)
    assert Zip.from_(s_ab).and_(one).all(Op.EQ)            # all(|x, y| x + y == 1.0)
    assert not Zip.from_(a).and_(b).all(Op.EQ)              # all(|x, y| x == y)
    assert not Zip.from_(s_ac).and_(one).all(Op.EQ)
    assert Zip.from_(s_ac).and_(one).any(Op.NE) and not Zip.from_(s_ab).and_(one).any(Op.NE)
    e0, e1 = Array.zeros(0, np.float32), Array.ones(0, np.float32)
    assert Zip.from_(e0).and_(e1).all(Op.EQ) and Zip.from_(e0).and_(e1).all(Op.NE)   # vacuously true
    assert not Zip.from_(e0).and_(e1).any(Op.EQ)
    x = np.arange(1000, dtype=np.int64)
    y = (x * 3 - 7).astype(np.int64)
    assert Zip.from_(Array.from_numpy(x)).and_(Array.from_numpy(y)).fold_sum(Op.MUL) == int((x * y).sum())
    with pytest.raises(nd.NdbError):
        Zip.from_(a).and_(b).all(Op.ADD)
    with pytest.raises(nd.ShapeError):
        Zip([a, Array.zeros(61, np.float32)]).all(Op.EQ)


EXACT_FNS = ["floor", "ceil", "round", "trunc", "fract", "signum", "recip", "pow2", "powi", "is_nan", "is_infinite"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_float_maths_maps(nd, oracle, dtype):
    """impl_float_maths.rs:80-160: exactly-rounded maps bit-exact, libm maps within a few ulp of the CPU libm."""
    from ndarray_b200.array import MATH_FN, MATH_FNS
    rng = np.random.default_rng(21)
    base = rng.uniform(-3, 3, 5000).astype(dtype)
    base[:8] = [0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 1.0]
    special = np.concatenate([base, np.array([np.inf, -np.inf, np.nan], dtype)])
    ulp = np.finfo(dtype).eps
    domain = {"ln": np.abs(base) + 1e-3, "log2": np.abs(base) + 1e-3, "log10": np.abs(base) + 1e-3, "ln_1p": np.abs(base),
              "asin": base / 3.01, "acos": base / 3.01, "acosh": np.abs(base) + 1.0, "atanh": base / 3.01,
              "log": np.abs(base) + 1e-3, "powf": np.abs(base) + 1e-3, "cbrt": base}
    param = {"powi": 3, "powf": 1.7, "log": 2.5, "abs_sub": 0.25, "hypot": 1.25}
    for name in MATH_FNS:
        x = special if name in EXACT_FNS else domain.get(name, base).astype(dtype)
        want = np.zeros_like(x)
        with np.errstate(all="ignore"):
            assert oracle.zip_(31, want, [x], [MATH_FN[name], param.get(name, 0.0)]) == 0
        dx = nd.Array.from_numpy(x)
        got = (getattr(dx, name)(param[name]) if name in param else getattr(dx, name)()).to_numpy()
        if name in EXACT_FNS:
            assert_bit_equal(got, want, name)
        else:
            scale = np.maximum(np.abs(want), np.finfo(dtype).tiny)
            rel = np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale
            assert np.nanmax(rel) <= 8 * ulp, (name, float(np.nanmax(rel)))
    c = nd.Array.from_numpy(special).clamp(-1.0, 2.0).to_numpy()
    wantc = np.zeros_like(special)
    assert oracle.zip_(32, wantc, [special], [-1.0, 2.0]) == 0
    assert_bit_equal(c, wantc, "clamp")
    assert np.array_equal(np.isnan(wantc), np.isnan(special))
    nanarr = nd.Array.from_numpy(np.array([np.nan, 1.0], dtype))
    assert nanarr.is_any_nan() and not nanarr.is_all_nan() and not nanarr.is_any_infinite()
    with pytest.raises(nd.NdbError):
        nd.Array.from_numpy(np.arange(4, dtype=np.int32)).map_math("exp")


def test_config5_full_size_64bit_indexing(nd):
    """BASELINE config 5 on one GPU: 2^32 + 40 f32 elements per array (64-bit indexing in every kernel class on the
    way), a = b*c + a in place followed by sum(); checked through constants, sampled slices and the exact count."""
    from ndarray_b200 import Array, Zip, Op, s_
    n = (1 << 32) + 40
    a, b, c = (Array.uninit((n,), np.float32) for _ in range(3))
    a.fill(0.25)
    b.fill(0.5)
    c.fill(1.5)
    Zip.from_(a).and_(b).and_(c).for_each("a=b*c+a")
    for lo in (0, (1 << 31) - 3, (1 << 32) - 5, n - 7):
        assert a.slice(s_[lo:lo + 7]).to_numpy().tolist() == [1.0] * 7
    total = float(a.sum())
    assert abs(total - n) <= 1e-6 * n
    # a strided 2-d view over the same buffer whose offsets exceed 2^32: rows kernel with 64-bit strides
    v = a.raw_view_with(8, (4096, 1 << 19), (1 << 20, 2))        # every other element of 4096 slabs 2^20 apart
    w = Array.zeros((4096, 1 << 19), np.float32)
    Zip.from_(w).and_(v).for_each("a=b")
    assert float(w.sum()) == 4096.0 * (1 << 19)
    v.fill(3.0)
    assert a.slice(s_[8:12]).to_numpy().tolist() == [3.0, 1.0, 3.0, 1.0]
    assert a.slice(s_[4095 * (1 << 20) + 8 + 2 * ((1 << 19) - 1) - 1:4095 * (1 << 20) + 8 + 2 * ((1 << 19) - 1) + 2]).to_numpy().tolist() == [1.0, 3.0, 1.0]
    # transposed view with a 2^31-element row pitch: tile kernel with 64-bit offsets
    t = a.raw_view_with(0, (2, 1 << 12), (1 << 31, 1)).t()            # shape (4096, 2), strides (1, 2^31)
    o = Array.zeros((1 << 12, 2), np.float32)
    Zip.from_(o).and_(t).for_each("a=b")
    got = o.to_numpy()
    assert got[:, 0].tolist()[:12] == [1.0] * 8 + [3.0, 1.0, 3.0, 1.0] and set(got[:, 1].tolist()) <= {1.0, 3.0}


def test_mapv_cast_matches_rust_as_semantics(nd, oracle):
    """`mapv(|x| x as B)` for every pair of element types on contiguous, transposed and reversed views."""
    rng = np.random.default_rng(31)
    src = {np.float32: rng.uniform(-3e9, 3e9, (37, 50)).astype(np.float32), np.float64: rng.uniform(-1e19, 1e19, (37, 50)),
           np.int32: rng.integers(-2 ** 31, 2 ** 31 - 1, (37, 50), dtype=np.int32),
           np.int64: rng.integers(-2 ** 63, 2 ** 63 - 1, (37, 50), dtype=np.int64)}
    for f in (np.float32, np.float64):
        src[f][0, :6] = [np.nan, np.inf, -np.inf, 0.5, -0.5, 2.0 ** 31]
    for ti, h in src.items():
        for fn in (lambda x: x, lambda x: x.T, lambda x: x[::-1, ::2]):
            m = Mirror(nd, h)
            hv, dv = m.view(fn)
            for to in (np.float32, np.float64, np.int32, np.int64):
                want = np.zeros(hv.shape, to)
                with np.errstate(all="ignore"):
                    assert oracle.cast(want, hv) == 0
                got = dv.mapv_cast(to).to_numpy()
                assert_bit_equal(got, want, f"{np.dtype(ti).name} as {np.dtype(to).name}")
    assert nd.Array.from_numpy(np.array([np.nan, 1e30, -1e30, -1.9], np.float32)).mapv_cast(np.int32).to_numpy().tolist() == \
        [0, 2 ** 31 - 1, -2 ** 31, -1]


HOST_CASES = {
    "contig": ((96, 40, 64), ident, [((96, 40, 64), ident), ((96, 40, 64), ident)]),
    "transposed": ((64, 40, 96), ident, [((96, 40, 64), lambda a: a.transpose(2, 1, 0)), ((64, 40, 96), ident)]),
    "neg_step_outer": ((48, 40, 64), ident, [((96, 40, 64), lambda a: a[::-1][::2]), ((96, 40, 64), lambda a: a[::2])]),
    "bcast_row": ((96, 40, 64), ident, [((96, 40, 64), ident), ((64,), lambda a: np.broadcast_to(a, (96, 40, 64)))]),
    "out_strided": ((96, 80), lambda a: a[:, ::2], [((96, 40), ident), ((96, 40), lambda a: a[::-1])]),
    "out_rows_stepped": ((96, 40), lambda a: a[::3], [((32, 40), ident), ((32, 40), ident)]),
    "one_d": ((100003,), ident, [((100003,), ident), ((100003,), lambda a: a[::-1])]),
    "empty": ((0, 5), ident, [((0, 5), ident), ((0, 5), ident)]),
}


@pytest.mark.parametrize("dtype", [np.float32, np.int64])
@pytest.mark.parametrize("case", sorted(HOST_CASES))
def test_zip_host_end_to_end_path(nd, oracle, case, dtype):
    """ndb_zip_host (the bench's e2e path): HOST views in, chunked H2D / kernel / D2H, result in the host buffer —
    bit-exact vs the oracle, and elements of the output buffer outside the output view untouched."""
    import ctypes as C
    from ndarray_b200._ffi import View, lib, check
    oshape, ofn, ins = HOST_CASES[case]
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    obuf = rand(rng, oshape, dtype)
    ibufs = [rand(rng, s, dtype) for s, _ in ins]
    want = obuf.copy()
    hv_want = ofn(want)
    hins = [f(b) for b, (_, f) in zip(ibufs, ins)]
    assert oracle.zip_(OP["ADD"], hv_want, hins) == 0

    def hview(a):
        v = View()
        v.ptr = a.__array_interface__["data"][0]
        v.ndim, v.dtype = a.ndim, {np.dtype(np.float32): 0, np.dtype(np.int64): 3}[a.dtype]
        for i in range(a.ndim):
            v.dim[i], v.stride[i] = a.shape[i], a.strides[i] // a.itemsize
        return v

    ctx = nd.default_context()
    parts = (View * 3)(hview(ofn(obuf)), hview(hins[0]), hview(hins[1]))
    check(lib().ndb_zip_host(ctx.handle, OP["ADD"], None, 3, parts))
    assert_bit_equal(obuf, want, case)
    # in-place: a += b through host buffers
    if case in ("contig", "out_strided", "one_d"):
        a0 = rand(rng, oshape, dtype)
        a1 = a0.copy()
        av, aw = ofn(a0), ofn(a1)
        assert oracle.zip_(OP["ADD"], aw, [aw, hins[1]]) == 0
        parts = (View * 3)(hview(av), hview(av), hview(hins[1]))
        check(lib().ndb_zip_host(ctx.handle, OP["ADD"], None, 3, parts))
        assert_bit_equal(a0, a1, case + " in place")


def test_zip_host_chunked_pipeline(nd, oracle):
    """Large enough (128 MiB per operand) that ndb_zip_host streams slabs: outermost-axis chunks, a stepped reversed
    operand moved with pitched copies, a transposed operand moved whole, and a non-dense output that is read back."""
    import ctypes as C
    from ndarray_b200._ffi import View, lib, check
    rng = np.random.default_rng(77)
    n0, n1, n2 = 256, 256, 512
    x = rng.uniform(-1, 1, (2 * n0, n1, n2)).astype(np.float32)
    yt = rng.uniform(-1, 1, (n2, n1, n0)).astype(np.float32)
    outbuf = rng.uniform(-1, 1, (n0, n1, 2 * n2)).astype(np.float32)
    want = outbuf.copy()
    xv, yv = x[::-1][::2], yt.transpose(2, 1, 0)
    assert oracle.zip_(OP["SUB"], want[:, :, ::2], [xv, yv], threads=8) == 0

    def hview(a):
        v = View()
        v.ptr = a.__array_interface__["data"][0]
        v.ndim, v.dtype = a.ndim, 0
        for i in range(a.ndim):
            v.dim[i], v.stride[i] = a.shape[i], a.strides[i] // a.itemsize
        return v

    parts = (View * 3)(hview(outbuf[:, :, ::2]), hview(xv), hview(yv))
    check(lib().ndb_zip_host(nd.default_context().handle, OP["SUB"], None, 3, parts))
    assert_bit_equal(outbuf, want, "chunked host path")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_fused_all_any_fold_paths_agree(nd, dtype):
    """Zip::all / any / fold on dense operands take the fused kernel, on strided ones the scratch path: same answers."""
    from ndarray_b200 import Array, Zip, Op
    rng = np.random.default_rng(41)
    n = 300007
    h1 = rand(rng, (n,), dtype)
    h2 = h1.copy()
    h2[n - 3] = h2[n - 3] + np.dtype(dtype).type(1)
    big1, big2 = np.zeros(2 * n, dtype), np.zeros(2 * n, dtype)
    big1[::2], big2[::2] = h1, h2
    for dense in (True, False):
        a = Array.from_numpy(h1) if dense else Array.from_numpy(big1).slice(nd.s_[::2])
        b = Array.from_numpy(h2) if dense else Array.from_numpy(big2).slice(nd.s_[::2])
        assert not Zip.from_(a).and_(b).all(Op.EQ) and Zip.from_(a).and_(a).all(Op.EQ)
        assert Zip.from_(a).and_(b).any(Op.NE) and not Zip.from_(a).and_(a).any(Op.NE)
        assert Zip.from_(a).and_(b).all(Op.LE) and not Zip.from_(a).and_(b).all(Op.GE)
        if np.dtype(dtype).kind == "i":
            with np.errstate(over="ignore"):
                assert Zip.from_(a).and_(b).fold_sum(Op.MUL) == (h1 * h2).sum(dtype=dtype)
                assert Zip.from_(a).and_(b).fold_sum(Op.SUB) == (h1 - h2).sum(dtype=dtype)
        else:
            want = float((h1.astype(np.float64) * h2.astype(np.float64)).sum())
            tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(n) * float(np.abs(h1 * h2).max())
            assert abs(float(Zip.from_(a).and_(b).fold_sum(Op.MUL)) - want) <= tol
