"""GPU parity: general_mat_mul / dot / general_mat_vec_mul and the CBLAS drop-in symbols.

The arithmetic of ndarray's GEMM lives in the un-vendored `matrixmultiply` crate, so — like ndarray's
own tests — parity is to tolerance against the naive and Kahan references (oracle modes 0 / 1):
  tests/oper.rs:595-635       relative 1e-7 / abs 1e-12 vs naive, strides {1,2,-1,-2}^2, 9 sizes, f64
  accuracy.rs:143-223         max|diff| / max|ref| < 1e-5 (f32), 1e-14 (f64; we allow 1e-12*sqrt(k)/k-scaled)
  tests/oper.rs:321-379       exact equality on integer-valued f32 across C/F operand layouts."""
import ctypes as C
import math

import numpy as np
import pytest

from .test_zip_gpu import Mirror

pytestmark = pytest.mark.gpu


def rel_norm(got, ref):
    return float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() / max(np.abs(ref).max(), 1e-300))


def nd_slice_np(a, s1, s2):
    """numpy view equal to ndarray's a.slice(s![..;s1, ..;s2]) (negative step = reversed, then strided)."""
    r = a[::-1] if s1 < 0 else a
    r = r[::abs(s1)]
    r = r[:, ::-1] if s2 < 0 else r
    return r[:, ::abs(s2)]


def range_mat(m, n, dtype):
    return np.arange(m * n, dtype=dtype).reshape(m, n)


def test_gen_mat_mul_reference_sweep(nd, oracle):  # tests/oper.rs:595-635
    alpha, beta = -2.3, math.pi
    sizes = [(4, 4, 4), (8, 8, 8), (17, 15, 16), (4, 17, 3), (17, 3, 22), (19, 18, 2), (16, 17, 15), (15, 16, 17), (67, 63, 62)]
    for s1 in (1, 2, -1, -2):
        for s2 in (1, 2, -1, -2):
            for (m, k, n) in sizes:
                ma, mb, mc = (Mirror(nd, range_mat(*shp, np.float64)) for shp in ((m, k), (k, n), (m, n)))
                ha, da = ma.view(lambda a: nd_slice_np(a, s1, s2))
                hb, db = mb.view(lambda a: nd_slice_np(a, s2, s2))
                hc, dc = mc.view(lambda a: nd_slice_np(a, s1, s2))
                assert oracle.gemm(alpha, ha, hb, beta, hc) == 0
                nd.general_mat_mul(alpha, da, db, beta, dc)
                np.testing.assert_allclose(mc.download(), mc.host, rtol=1e-7, atol=1e-12)


def test_gen_mat_mul_i32_exact(nd, oracle):  # tests/oper.rs:653-681
    for (m, k, n) in [(4, 4, 4), (17, 15, 16), (67, 63, 62)]:
        a, b, c = range_mat(m, k, np.int32), range_mat(k, n, np.int32), range_mat(m, n, np.int32)
        want = c.copy()
        assert oracle.gemm(-1, a, b, 1, want) == 0
        dc = nd.Array.from_numpy(c)
        nd.general_mat_mul(-1, nd.Array.from_numpy(a), nd.Array.from_numpy(b), 1, dc)
        assert np.array_equal(dc.to_numpy(), want)


def test_mat_mul_layouts_exact(nd):  # tests/oper.rs:321-399
    for (m, n, k) in [(8, 8, 8), (10, 17, 9), (64, 64, 64), (130, 70, 40)]:
        a = range_mat(m, n, np.float32) % 16
        b = (range_mat(n, k, np.float32) % 8) / np.float32(4.)
        want = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
        for af in (False, True):
            for bf in (False, True):
                da = nd.Array.from_numpy(np.asfortranarray(a) if af else a)
                db = nd.Array.from_numpy(np.asfortranarray(b) if bf else b)
                c = da.dot(db)
                assert np.array_equal(c.to_numpy(), want)
                # dot's output is F-order only when both inputs have unit row stride (impl_linalg.rs:319-327)
                assert (c.strides == (1, m)) == (af and bf and m > 1 and k > 1) or m == 1 or k == 1
                # transposed product equals product of transposes
                ct = db.t().dot(da.t())
                assert np.array_equal(ct.to_numpy(), want.T)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-14)])
def test_accuracy_random(nd, oracle, dtype, tol):  # accuracy.rs:143-223
    rng = np.random.default_rng(42)
    for _ in range(6):
        m, k, n = int(rng.integers(15, 128)), int(rng.integers(15, 128)), int(rng.integers(15, 512))
        a = rng.standard_normal((m, k)).astype(dtype)
        b = rng.standard_normal((k, n)).astype(dtype)
        ref = np.zeros((m, n), dtype)
        assert oracle.gemm(1, a, b, 0, ref, mode=1, threads=4) == 0
        c = nd.Array.from_numpy(a).dot(nd.Array.from_numpy(b)).to_numpy()
        assert rel_norm(c, ref) < tol * (1 if dtype == np.float32 else 8)
        # stride-2 operands (accuracy.rs:205-223)
        a2 = rng.standard_normal((2 * m, 2 * k)).astype(dtype)
        ma = Mirror(nd, a2)
        ha, da = ma.view(lambda x: x[::2, ::2])
        ref2 = np.zeros((m, n), dtype)
        oracle.gemm(1, ha, b, 0, ref2, mode=1, threads=4)
        c2 = da.dot(nd.Array.from_numpy(b)).to_numpy()
        assert rel_norm(c2, ref2) < tol * (1 if dtype == np.float32 else 8)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_medium_gemm_all_layouts(nd, dtype):
    """512..1024-sized problems that take the tensor-core path, every C/F combination + alpha/beta, against
    numpy in float64."""
    rng = np.random.default_rng(3)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    for (m, k, n) in [(512, 512, 512), (768, 1024, 640), (1000, 520, 776), (129, 2048, 257)]:
        a = rng.standard_normal((m, k)).astype(dtype)
        b = rng.standard_normal((k, n)).astype(dtype)
        c0 = rng.standard_normal((m, n)).astype(dtype)
        for alpha, beta in [(1.0, 0.0), (-2.3, math.pi)]:
            ref = alpha * (a.astype(np.float64) @ b.astype(np.float64)) + beta * c0.astype(np.float64)
            for af in (False, True):
                for bf in (False, True):
                    for cf in (False, True):
                        da = nd.Array.from_numpy(np.asfortranarray(a) if af else a)
                        db = nd.Array.from_numpy(np.asfortranarray(b) if bf else b)
                        dc = nd.Array.from_numpy(np.asfortranarray(c0) if cf else c0)
                        nd.general_mat_mul(alpha, da, db, beta, dc)
                        err = rel_norm(dc.to_numpy(), ref)
                        assert err < tol * math.sqrt(k) / 8, (m, k, n, alpha, af, bf, cf, err)


def test_beta_zero_never_reads_c(nd):
    a = nd.Array.from_numpy(np.eye(64, dtype=np.float64))
    b = nd.Array.from_numpy(np.full((64, 64), 2.0))
    c = nd.Array.from_numpy(np.full((64, 64), np.nan))
    nd.general_mat_mul(1.0, a, b, 0.0, c)
    assert np.array_equal(c.to_numpy(), np.full((64, 64), 2.0))
    big = nd.Array.from_numpy(np.full((512, 512), np.nan))
    nd.general_mat_mul(1.0, nd.Array.from_numpy(np.eye(512)), nd.Array.from_numpy(np.full((512, 512), 3.0)), 0.0, big)
    assert np.array_equal(big.to_numpy(), np.full((512, 512), 3.0))


def test_zero_sized_and_mismatch(nd):  # tests/oper.rs:464-483
    for (m, k, n) in [(0, 4, 5), (4, 0, 5), (4, 5, 0)]:
        a = nd.Array.zeros((m, k), np.float64)
        b = nd.Array.zeros((k, n), np.float64)
        c = a.dot(b)
        assert c.shape == (m, n)
        assert np.array_equal(c.to_numpy(), np.zeros((m, n)))
    with pytest.raises(nd.ShapeError, match="not compatible for matrix multiplication"):
        nd.general_mat_mul(1.0, nd.Array.zeros((3, 4), np.float64), nd.Array.zeros((5, 3), np.float64), 0.0,
                           nd.Array.zeros((3, 3), np.float64))


def test_broadcast_and_reversed_operands(nd):  # tests/oper.rs:427-460
    rng = np.random.default_rng(1)
    a = rng.standard_normal((8, 8))
    x = rng.standard_normal(8)
    da = nd.Array.from_numpy(a)
    bx = nd.Array.from_numpy(x).broadcast((8, 8))
    np.testing.assert_allclose(da.dot(bx).to_numpy(), a @ np.broadcast_to(x, (8, 8)), rtol=1e-12)
    rev = da.slice(nd.s_[::-1, ::-1])
    np.testing.assert_allclose(rev.dot(da).to_numpy(), a[::-1, ::-1] @ a, rtol=1e-12)


def test_gemv_and_vec_mat(nd, oracle):  # tests/oper.rs:686-790
    """general_mat_vec_mul / mat.dot(vec) / vec.dot(mat) against the ORACLE: ndarray computes a gemv row by row with its
    1-D dot (impl_linalg.rs:657-743 -> numeric_util.rs:56-93 unrolled_dot), so the reference for y[i] is
    alpha * orc_dot(a[i, :], x) + beta * y[i] in the element type; tolerance 1e-12 (f64) / 1e-5 (f32) x sqrt(k)."""
    rng = np.random.default_rng(2)
    for dt, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for (m, k) in [(10, 10), (17, 15), (64, 130), (300, 77)]:
            a = rng.standard_normal((m, k)).astype(dt)
            x = rng.standard_normal(k).astype(dt)
            y0 = rng.standard_normal(m).astype(dt)
            v = rng.standard_normal(m).astype(dt)
            ax = np.array([oracle.dot(np.ascontiguousarray(a[i]), x) for i in range(m)], dtype=dt)
            va = np.array([oracle.dot(v, np.ascontiguousarray(a[:, j])) for j in range(k)], dtype=dt)
            want_y = dt(-2.3) * ax + dt(math.pi) * y0
            lim = tol * np.sqrt(k) * max(1.0, np.abs(want_y).max())
            for order in ("C", "F"):
                da = nd.Array.from_numpy(np.asfortranarray(a) if order == "F" else a)
                y = nd.Array.from_numpy(y0)
                nd.general_mat_vec_mul(-2.3, da, nd.Array.from_numpy(x), math.pi, y)
                assert np.abs(y.to_numpy() - want_y).max() <= lim, (dt, m, k, order)
                assert np.abs(da.dot(nd.Array.from_numpy(x)).to_numpy() - ax).max() <= lim
                assert np.abs(nd.Array.from_numpy(v).dot(da).to_numpy() - va).max() <= tol * np.sqrt(m) * max(1.0, np.abs(va).max())


def test_dot_1d_against_the_oracle(nd, oracle):
    """1-D dot (impl_linalg.rs:74-133): the oracle restates unrolled_dot's 8-lane association; integers are exact,
    floats agree within 1e-12 / 1e-5 x sqrt(n) of the sum of |products|/n (summation order differs)."""
    rng = np.random.default_rng(12)
    for n in (0, 1, 7, 8, 1000, 100003):
        for dt in (np.float64, np.float32, np.int32, np.int64):
            if np.dtype(dt).kind == "i":
                x = rng.integers(-2 ** 20, 2 ** 20, n).astype(dt)
                y = rng.integers(-2 ** 20, 2 ** 20, n).astype(dt)
            else:
                x, y = rng.standard_normal(n).astype(dt), rng.standard_normal(n).astype(dt)
            dx, dy = nd.Array.from_numpy(x), nd.Array.from_numpy(y)
            cases = [(x, y, dx, dy), (x[::-1], y, dx.slice(nd.s_[::-1]), dy)]
            if n > 1:
                h = len(x[::2])
                cases.append((x[::2], y[1::2] if len(y[1::2]) == h else y[:h], dx.slice(nd.s_[::2]),
                              dy.slice(nd.s_[1::2]) if len(y[1::2]) == h else dy.slice(nd.s_[:h])))
            for hx, hy, vx, vy in cases:
                got = vx.dot(vy)
                want = oracle.dot(np.ascontiguousarray(hx), np.ascontiguousarray(hy))
                if np.dtype(dt).kind == "i":
                    assert got == want, (n, dt)
                else:
                    tol = (1e-12 if dt == np.float64 else 1e-5) * np.sqrt(max(n, 1)) * max(1e-30, float(np.abs(hx * hy).mean()) if n else 1.0)
                    assert abs(float(got) - float(want)) <= tol, (n, dt, got, want)


def test_cblas_dropin_symbols(nd):
    """Call the exported cblas_dgemm / cblas_sgemm / cblas_ddot with HOST pointers, the way ndarray's `blas`
    feature does (impl_linalg.rs:466-491): row-major C with transposed operands, ld > n padding kept intact."""
    L = nd.lib()
    L.cblas_dgemm.restype = None
    L.cblas_dgemm.argtypes = [C.c_int] * 6 + [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    L.cblas_sgemm.restype = None
    L.cblas_sgemm.argtypes = [C.c_int] * 6 + [C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
    L.cblas_ddot.restype = C.c_double
    L.cblas_ddot.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(8)
    m, n, k = 33, 21, 47
    a = rng.standard_normal((m, k))
    bt = rng.standard_normal((n, k + 3))          # B stored transposed with padding: ldb = k + 3
    c = rng.standard_normal((m, n + 5))           # ldc = n + 5
    c0 = c.copy()
    L.cblas_dgemm(101, 111, 112, m, n, k, -2.3, a.ctypes.data, k, bt.ctypes.data, k + 3, 0.5, c.ctypes.data, n + 5)
    np.testing.assert_allclose(c[:, :n], -2.3 * a @ bt[:, :k].T + 0.5 * c0[:, :n], rtol=1e-12)
    assert np.array_equal(c[:, n:], c0[:, n:])    # padding untouched
    af = np.asfortranarray(a.astype(np.float32))
    bf = np.asfortranarray(rng.standard_normal((k, n)).astype(np.float32))
    cf = np.zeros((m, n), np.float32, order="F")
    L.cblas_sgemm(102, 111, 111, m, n, k, 1.0, af.ctypes.data, m, bf.ctypes.data, k, 0.0, cf.ctypes.data, m)
    assert rel_norm(cf, af.astype(np.float64) @ bf.astype(np.float64)) < 1e-5
    x, y = rng.standard_normal(100), rng.standard_normal(100)
    got = L.cblas_ddot(50, x.ctypes.data, 2, y.ctypes.data, -2)
    assert abs(got - float(x[:100:2] @ y[:100:2][::-1])) < 1e-12


def test_cblas_gemv_dropin(nd):  # impl_linalg.rs:705-728
    L = nd.lib()
    L.cblas_dgemv.restype = None
    L.cblas_dgemv.argtypes = [C.c_int] * 4 + [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    L.cblas_sgemv.restype = None
    L.cblas_sgemv.argtypes = [C.c_int] * 4 + [C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
    rng = np.random.default_rng(4)
    m, n = 37, 53
    a = rng.standard_normal((m, n + 2))            # row-major, lda = n + 2
    x, y = rng.standard_normal(2 * n), rng.standard_normal(m)
    y0 = y.copy()
    L.cblas_dgemv(101, 111, m, n, -2.3, a.ctypes.data, n + 2, x.ctypes.data, 2, 0.5, y.ctypes.data, 1)
    np.testing.assert_allclose(y, -2.3 * a[:, :n] @ x[::2] + 0.5 * y0, rtol=1e-12)
    xt, yt = rng.standard_normal(m), rng.standard_normal(2 * n)
    yt0 = yt.copy()
    L.cblas_dgemv(101, 112, m, n, 1.0, a.ctypes.data, n + 2, xt.ctypes.data, 1, 1.0, yt.ctypes.data, -2)   # y += A^T x, y reversed
    np.testing.assert_allclose(yt[::2][::-1], a[:, :n].T @ xt + yt0[::2][::-1], rtol=1e-12)
    assert np.array_equal(yt[1::2], yt0[1::2])      # gaps untouched
    af = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
    xs, ys = rng.standard_normal(n).astype(np.float32), np.zeros(m, np.float32)
    L.cblas_sgemv(102, 111, m, n, 1.0, af.ctypes.data, m, xs.ctypes.data, 1, 0.0, ys.ctypes.data, 1)
    np.testing.assert_allclose(ys, af.astype(np.float64) @ xs.astype(np.float64), rtol=2e-5, atol=1e-5)


def test_large_gemv_both_layouts(nd):
    rng = np.random.default_rng(5)
    for dtype, tol in ((np.float32, 2e-5), (np.float64, 1e-12)):
        a = rng.standard_normal((5000, 3000)).astype(dtype)
        x = rng.standard_normal(3000).astype(dtype)
        ref = a.astype(np.float64) @ x.astype(np.float64)
        for arr in (a, np.asfortranarray(a)):
            y = nd.Array.from_numpy(arr).dot(nd.Array.from_numpy(x)).to_numpy()
            assert np.abs(y - ref).max() / np.abs(ref).max() < tol * math.sqrt(3000) / 8


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-13)])
def test_strided_operands_take_the_tensor_path(nd, oracle, dtype, tol):
    """Stride-2 / reversed / broadcast operands and a strided, reversed C at a size the tensor-core kernels accept
    (they pack such operands through the Zip engine): values vs numpy float64, and untouched gaps in C."""
    rng = np.random.default_rng(17)
    m, k, n = 200, 136, 264
    alpha, beta = -2.3, math.pi
    for (sa, sb, sc) in [((2, -1), (-2, 2), (-1, 2)), ((-2, 2), (1, -1), (2, -2)), ((1, 1), (2, 1), (1, -1))]:
        ma = Mirror(nd, rng.standard_normal((m * abs(sa[0]), k * abs(sa[1]))).astype(dtype))
        mb = Mirror(nd, rng.standard_normal((k * abs(sb[0]), n * abs(sb[1]))).astype(dtype))
        mc = Mirror(nd, rng.standard_normal((m * abs(sc[0]), n * abs(sc[1]))).astype(dtype))
        c0 = mc.host.copy()
        ha, da = ma.view(lambda x: nd_slice_np(x, *sa))
        hb, db = mb.view(lambda x: nd_slice_np(x, *sb))
        hc, dc = mc.view(lambda x: nd_slice_np(x, *sc))
        want = alpha * (ha.astype(np.float64) @ hb.astype(np.float64)) + beta * hc.astype(np.float64)
        nd.general_mat_mul(alpha, da, db, beta, dc)
        got_full = mc.download()
        got = nd_slice_np(got_full, *sc)
        assert rel_norm(got, want) < tol * math.sqrt(k), (sa, sb, sc)
        mask = np.ones_like(c0, dtype=bool)
        nd_slice_np(mask, *sc)[...] = False
        assert np.array_equal(got_full[mask], c0[mask]), "elements outside the C view were modified"
    # broadcast B (a row repeated): tests/oper.rs:427-445
    a = rng.standard_normal((m, k)).astype(dtype)
    row = rng.standard_normal(n).astype(dtype)
    bb = nd.Array.from_numpy(row).broadcast((k, n))
    c = nd.Array.from_numpy(a).dot(bb).to_numpy()
    assert rel_norm(c, a.astype(np.float64) @ np.broadcast_to(row.astype(np.float64), (k, n))) < tol * math.sqrt(k)


def test_f64_tma_and_cpasync_kernels_agree_bitwise(nd, monkeypatch):
    """The TMA-fed and the cp.async-fed f64 kernels run the same DMMA sequence per output element, so they must agree to
    the bit on every layout combination, with tile overhang (m, n not multiples of 128) and a partial last k-tile."""
    rng = np.random.default_rng(11)
    monkeypatch.setenv("NDB_GEMM_F64_SPLITK", "1")  # split-K regroups the k sum (covered to tolerance below)
    for (m, k, n) in [(256, 256, 256), (1000, 520, 776), (132, 36, 388)]:
        a = rng.standard_normal((m, k))
        b = rng.standard_normal((k, n))
        c0 = rng.standard_normal((m, n))
        for af in (False, True):
            for bf in (False, True):
                for cf in (False, True):
                    outs = []
                    for force_cpasync in (False, True):
                        if force_cpasync:
                            monkeypatch.setenv("NDB_GEMM_F64_CPASYNC", "1")
                        else:
                            monkeypatch.delenv("NDB_GEMM_F64_CPASYNC", raising=False)
                        da = nd.Array.from_numpy(np.asfortranarray(a) if af else a)
                        db = nd.Array.from_numpy(np.asfortranarray(b) if bf else b)
                        dc = nd.Array.from_numpy(np.asfortranarray(c0) if cf else c0)
                        nd.general_mat_mul(-2.3, da, db, math.pi, dc)
                        outs.append(dc.to_numpy())
                    monkeypatch.delenv("NDB_GEMM_F64_CPASYNC", raising=False)
                    assert np.array_equal(outs[0], outs[1]), (m, k, n, af, bf, cf)
                    ref = -2.3 * (a @ b) + math.pi * c0
                    assert rel_norm(outs[0], ref) < 1e-12 * math.sqrt(k)


def test_f64_small_outputs_split_k_over_clusters(nd, monkeypatch):
    """Small outputs use 64x64 tiles and split K over a thread-block cluster (partials merged through distributed shared
    memory in rank order): every cluster size gives the one-CTA result within rounding, and repeated runs are
    bit-identical (fixed merge order)."""
    rng = np.random.default_rng(12)
    for (m, k, n) in [(256, 256, 256), (512, 512, 512), (192, 1024, 320), (100, 264, 60)]:
        a = rng.standard_normal((m, k))
        b = rng.standard_normal((k, n))
        c0 = rng.standard_normal((m, n))
        ref = -2.3 * (a @ b) + math.pi * c0
        for bf in (False, True):
            res = {}
            for split in ("1", "2", "4", "8", None):
                if split is None:
                    monkeypatch.delenv("NDB_GEMM_F64_SPLITK", raising=False)
                else:
                    monkeypatch.setenv("NDB_GEMM_F64_SPLITK", split)
                runs = []
                for _ in range(2):
                    da = nd.Array.from_numpy(a)
                    db = nd.Array.from_numpy(np.asfortranarray(b) if bf else b)
                    dc = nd.Array.from_numpy(c0)
                    nd.general_mat_mul(-2.3, da, db, math.pi, dc)
                    runs.append(dc.to_numpy())
                assert np.array_equal(runs[0], runs[1]), (m, k, n, split)
                res[split] = runs[0]
                assert rel_norm(runs[0], ref) < 1e-12 * math.sqrt(k), (m, k, n, split)
            monkeypatch.delenv("NDB_GEMM_F64_SPLITK", raising=False)


def test_f64_stream_k_matches_whole_tile_kernel(nd, monkeypatch):
    """Stream-K (gemm_f64_tma_streamk: k-ranges cut across tile boundaries, partials handed over through global memory in
    ticket order) against the one-tile-per-CTA kernel: same result within rounding on every layout combination, with tile
    overhang, with alpha/beta, bit-identical between runs.  Shapes cover: the automatic case (more tiles than SMs with a
    ragged last wave), a forced small case where one tile is shared by several CTAs (64 tiles over 148 CTAs), and data-
    parallel tiles after the stream-K region (3 waves and a remainder)."""
    rng = np.random.default_rng(21)
    monkeypatch.setenv("NDB_GEMM_F64_TILE", "128")
    for (m, k, n) in [(2048, 512, 2048), (1024, 1024, 1024), (1000, 520, 776), (3000, 256, 3100), (1664, 288, 1700), (640, 2048, 384)]:
        a = rng.standard_normal((m, k))
        b = rng.standard_normal((k, n))
        c0 = rng.standard_normal((m, n))
        ref = -2.3 * (a @ b) + math.pi * c0
        for af, bf in ((False, False), (True, False), (False, True), (True, True)):
            res = {}
            for sk in ("0", "1"):
                monkeypatch.setenv("NDB_GEMM_F64_STREAMK", sk)
                runs = []
                for _ in range(2):
                    da = nd.Array.from_numpy(np.asfortranarray(a) if af else a)
                    db = nd.Array.from_numpy(np.asfortranarray(b) if bf else b)
                    dc = nd.Array.from_numpy(c0)
                    nd.general_mat_mul(-2.3, da, db, math.pi, dc)
                    runs.append(dc.to_numpy())
                assert np.array_equal(runs[0], runs[1]), (m, k, n, af, bf, sk)
                assert rel_norm(runs[0], ref) < 1e-12 * math.sqrt(k), (m, k, n, af, bf, sk)
                res[sk] = runs[0]
            assert rel_norm(res["1"], res["0"]) < 1e-14 * math.sqrt(k), (m, k, n, af, bf)
    # beta == 0 never reads C on the stream-K path either (impl_linalg.rs:323-329)
    monkeypatch.setenv("NDB_GEMM_F64_STREAMK", "1")
    a = rng.standard_normal((1024, 1024))
    b = rng.standard_normal((1024, 1024))
    dc = nd.Array.from_numpy(np.full((1024, 1024), np.nan))
    nd.general_mat_mul(1.0, nd.Array.from_numpy(a), nd.Array.from_numpy(b), 0.0, dc)
    assert rel_norm(dc.to_numpy(), a @ b) < 1e-12 * 32
    monkeypatch.delenv("NDB_GEMM_F64_STREAMK", raising=False)
    monkeypatch.delenv("NDB_GEMM_F64_TILE", raising=False)
    # the automatic choice: 2048^2 outputs are 256 tiles on 148 SMs
    a = rng.standard_normal((2048, 1024))
    b = rng.standard_normal((1024, 2048))
    got = nd.Array.from_numpy(a).dot(nd.Array.from_numpy(b)).to_numpy()
    assert rel_norm(got, a @ b) < 1e-12 * 32


def test_f64_stream_k_replays_inside_a_captured_graph(nd, monkeypatch):
    """The stream-K hand-over state (ticket, flags) returns to zero when the grid retires, so a captured launch can be
    replayed: three replays over changing inputs, each checked."""
    rng = np.random.default_rng(22)
    ctx = nd.default_context()
    n = 2048
    a = rng.standard_normal((n, 256))
    b = rng.standard_normal((256, n))
    da, db = nd.Array.from_numpy(a), nd.Array.from_numpy(b)
    dc = nd.Array.zeros((n, n), np.float64)
    nd.general_mat_mul(1.0, da, db, 0.0, dc)  # first call outside the capture: allocations, attributes
    with ctx.capture() as g:
        nd.general_mat_mul(1.0, da, db, 0.0, dc)
    for scale in (1.0, -3.0, 0.5):
        da.assign(nd.Array.from_numpy(a * scale))
        g.launch()
        ctx.sync()
        assert rel_norm(dc.to_numpy(), (a * scale) @ b) < 1e-12 * 16, scale


def test_cblas_large_operands_bounce_pipeline(nd):
    """Operands above 4 MiB travel through the pinned bounce-buffer pipeline (several 16 MiB chunks each way, pageable
    numpy memory like ndarray's Vec): results and the ldc padding must survive the trip."""
    L = nd.lib()
    L.cblas_dgemm.restype = None
    L.cblas_dgemm.argtypes = [C.c_int] * 6 + [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    rng = np.random.default_rng(21)
    m, n, k = 2304, 2048, 1536
    a = rng.standard_normal((m, k))               # 27 MiB
    b = rng.standard_normal((k, n))               # 24 MiB
    c = rng.standard_normal((m, n + 8))           # 36 MiB, ldc = n + 8
    c0 = c.copy()
    L.cblas_dgemm(101, 111, 111, m, n, k, 1.5, a.ctypes.data, k, b.ctypes.data, n, -0.5, c.ctypes.data, n + 8)
    ref = 1.5 * (a @ b) - 0.5 * c0[:, :n]
    assert rel_norm(c[:, :n], ref) < 1e-12 * math.sqrt(k)
    assert np.array_equal(c[:, n:], c0[:, n:])


def test_cblas_pipelined_row_blocks(nd):
    """Large row-major products with an untransposed A take the pipelined path (B up whole, A and C in row blocks, a worker
    thread draining C): every layout ndarray's `blas` feature produces must come back equal to the one-shot path's
    result (to rounding) and within tolerance of the reference - row-major, column-major (exchanged into row-major inside), transposed B,
    padded ldc with beta != 0, and a row count that is not a multiple of the block."""
    import os
    L = nd.lib()
    L.cblas_dgemm.restype = None
    L.cblas_dgemm.argtypes = [C.c_int] * 6 + [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    L.cblas_sgemm.restype = None
    L.cblas_sgemm.argtypes = [C.c_int] * 6 + [C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
    rng = np.random.default_rng(33)
    m, n, k = 4200, 1030, 2100                      # A: 67 MiB (f64) -> pipelined; 4200 rows = 2 blocks of 2176 / 2024
    a = rng.standard_normal((m, k))
    b = rng.standard_normal((k, n))
    bt = np.ascontiguousarray(b.T)
    c0 = rng.standard_normal((m, n + 6))
    ref = -2.3 * (a @ b) + math.pi * c0[:, :n]

    def run(pipelined, transb):
        if pipelined:
            os.environ.pop("NDB_CBLAS_NO_PIPELINE", None)
        else:
            os.environ["NDB_CBLAS_NO_PIPELINE"] = "1"
        c = c0.copy()
        if transb:
            L.cblas_dgemm(101, 111, 112, m, n, k, -2.3, a.ctypes.data, k, bt.ctypes.data, k, math.pi, c.ctypes.data, n + 6)
        else:
            L.cblas_dgemm(101, 111, 111, m, n, k, -2.3, a.ctypes.data, k, b.ctypes.data, n, math.pi, c.ctypes.data, n + 6)
        os.environ.pop("NDB_CBLAS_NO_PIPELINE", None)
        return c

    for transb in (False, True):
        one, pipe = run(False, transb), run(True, transb)
        assert rel_norm(pipe[:, :n], ref) < 1e-12 * math.sqrt(k)
        assert np.array_equal(pipe[:, n:], c0[:, n:])          # padding untouched
        # a block's tile count decides where stream-K cuts the k sum, so the two paths agree to rounding, not to the bit
        assert rel_norm(one[:, :n], pipe[:, :n]) < 1e-14 * math.sqrt(k)
    # column-major C = A B with A (m x k) stored column-major is the row-major product C^T = B^T A^T: pipelined over B's columns
    af, bf = np.asfortranarray(a[:1100]), np.asfortranarray(rng.standard_normal((k, 4300)))
    cf = np.zeros((1100, 4300), order="F")
    L.cblas_dgemm(102, 111, 111, 1100, 4300, k, 1.0, af.ctypes.data, 1100, bf.ctypes.data, k, 0.0, cf.ctypes.data, 1100)
    assert rel_norm(cf, af @ bf) < 1e-12 * math.sqrt(k)
    # f32
    a32, b32 = rng.standard_normal((8300, 2100)).astype(np.float32), rng.standard_normal((2100, 520)).astype(np.float32)
    c32 = np.zeros((8300, 520), np.float32)
    L.cblas_sgemm(101, 111, 111, 8300, 520, 2100, 1.0, a32.ctypes.data, 2100, b32.ctypes.data, 520, 0.0, c32.ctypes.data, 520)
    assert rel_norm(c32, a32.astype(np.float64) @ b32.astype(np.float64)) < 1e-5
