"""GPU parity for the rayon-facing surface (`par_azip!`, `Zip::par_*`, parallel iterators): ndarray's own parallel tests
(tests/par_azip.rs, tests/par_zip.rs, tests/par_rayon.rs) run at THEIR sizes (M = 10240, N = 100) through the same entry
points as the sequential forms - on the GPU the grid is the rayon pool (src/parallel/impl_par_methods.rs:64), so `par_*`
must give exactly what the sequential call gives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, N = 1024 * 10, 100


def test_par_azip_reference_cases(nd):  # tests/par_azip.rs:9-96
    from ndarray_b200 import Array, Zip, par_azip
    a = Array.zeros((62,), np.int64)                                           # test_par_azip1
    a.fill(42)
    assert np.array_equal(a.to_numpy(), Array.from_elem((62,), 42, np.int64).to_numpy())
    i, j = np.indices((5, 7))                                                  # test_par_azip2
    with np.errstate(divide="ignore"):
        hb = (1.0 / (i + 2 * j).astype(np.float32)).astype(np.float32)          # 1/0 = inf at (0, 0), as in Rust
    a2, b2 = Array.zeros((5, 7), np.float32), Array.from_numpy(hb)
    par_azip("a=b", a2, b2)
    assert a2.to_numpy().tobytes() == hb.tobytes()
    n = 62                                                                     # test_par_azip9: nine producers
    ins = [Array.from_numpy((np.arange(n) * (1 << k)).astype(np.int32)) for k in range(8)]
    out = Array.zeros((n,), np.int32)
    par_azip("a=sum", out, *ins)
    assert np.array_equal(out.to_numpy(), (np.arange(n) * 255).astype(np.int32))
    with pytest.raises(nd.NdbError, match="Producer dimension mismatch"):       # test_zip_dim_mismatch_1
        Zip.from_(Array.zeros((5, 7), np.float32)).and_(Array.zeros((6, 7), np.float32))
    a1 = Array.from_numpy(np.arange(12, dtype=np.int64))                        # test_indices_1
    z = Zip.indexed(a1)
    seen = Array.zeros((12,), np.int64)
    Zip.from_(seen).and_(z.parts[1]).par_for_each("a=b")
    assert len(z.parts) == 2 and np.array_equal(seen.to_numpy(), a1.to_numpy())


def test_par_zip_indexed_reference_shapes(nd):  # tests/par_zip.rs:19-78
    from ndarray_b200 import Array, Zip
    for shape in ((10, 10), (M, N), (1, 2, 1, 2, 3)):
        out = [Array.zeros(shape, np.int64) for _ in shape]
        z = Zip.indexed(out[0])
        assert len(z.parts) == 1 + len(shape)
        for k in range(len(shape)):
            Zip.from_(out[k]).and_(z.parts[1 + k]).par_for_each("a=b")
        want = np.indices(shape)
        for k in range(len(shape)):
            assert np.array_equal(out[k].to_numpy(), want[k]), (shape, k)


def test_par_zip_map_and_collect(nd):  # tests/par_zip.rs:11-17, :80-159
    from ndarray_b200 import Array, Op, Zip
    a = Array.zeros((M, N), np.float64)                                        # test_zip_1: x = x.exp()
    assert np.array_equal(a.exp().to_numpy(), np.ones((M, N)))
    i, j = np.indices((M, N))
    hb = (1.0 / (i + 2 * j + 1).astype(np.float32)).astype(np.float32)
    hc = np.log((1 + i + j).astype(np.float32)).astype(np.float32)
    b, c = Array.from_numpy(hb), Array.from_numpy(hc)
    r = Zip.from_(b).and_(c).par_map_collect(Op.ADD)                           # test_zip_collect
    assert r.to_numpy().tobytes() == (hb + hc).tobytes() and r.strides == b.strides
    rt = Zip.from_(b.t()).and_(c.t()).par_map_collect(Op.ADD)
    assert rt.to_numpy().tobytes() == (hb + hc).T.tobytes() and rt.strides == b.t().strides
    into = Array.zeros((M, N), np.float32)                                     # test_zip_assign_into
    Zip.from_(b).and_(c).par_map_assign_into(into, Op.ADD)
    assert into.to_numpy().tobytes() == (hb + hc).tobytes()
    for m in range(1, 32, 5):                                                  # test_zip_small_collect (a stride through its 0..32 x 0..16 grid)
        for n in range(1, 16, 3):
            for order in ("C", "F"):
                i, j = np.indices((m, n))
                sb = np.asarray(1.0 / (i + 2 * j + 1), dtype=np.float32, order=order)
                sc = np.asarray(np.log((1 + i + j).astype(np.float32)), dtype=np.float32, order=order)
                db, dc = Array.from_numpy(sb), Array.from_numpy(sc)
                r = Zip.from_(db).and_(dc).par_map_collect(Op.ADD)
                assert r.to_numpy().tobytes() == (sb + sc).tobytes(order="C"), (m, n, order)
                if m > 1 and n > 1:
                    assert r.strides == db.strides, (m, n, order)


def test_par_iter_sums_equal_the_plain_sum(nd):  # tests/par_rayon.rs:12-24, :44-56
    from ndarray_b200 import Array, Axis
    h = np.repeat(np.arange(M, dtype=np.float64)[:, None], N, axis=1)           # row i filled with i
    a = Array.from_numpy(h)
    total = a.sum()
    assert total == h.sum() == N * M * (M - 1) / 2                              # integers below 2^53: every order is exact
    assert a.sum_axis(Axis(1)).sum() == total                                   # axis_iter().into_par_iter().map(sum).sum()
    assert a.sum_axis(Axis(0)).sum() == total
