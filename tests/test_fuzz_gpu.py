"""Seeded random-geometry parity: the planner/kernels against the oracle on views no hand-written case list covers.

Each case draws a common logical shape (rank 1..5) and, per producer, a random route to it: an oversized base buffer,
slices with steps in {-3..3}\\{0}, a random axis permutation, inverted axes, and (inputs only) axes broadcast from
length 1.  The output view never aliases itself (no broadcast on the output).  Elementwise results are compared
bit-for-bit; reductions along a random axis exactly for integers and to the north-star tolerance for floats."""
import os

import numpy as np
import pytest

from .conftest import assert_bit_equal
from .test_zip_gpu import Mirror, rand

pytestmark = pytest.mark.gpu


def random_view_fn(rng, shape, allow_broadcast):
    """Returns (base_shape, fn) with fn(base).shape == shape."""
    nd = len(shape)
    perm = rng.permutation(nd)
    steps, offs, bcast = [], [], []
    base_shape = [0] * nd
    for ax in range(nd):
        n = shape[ax]
        b = allow_broadcast and n > 1 and rng.random() < 0.2
        step = int(rng.choice([1, 1, 1, -1, 2, -2, 3, -3]))
        off = int(rng.integers(0, 3))
        tail = int(rng.integers(0, 3))
        bcast.append(b)
        steps.append(step)
        offs.append(off)
        base_len = 1 if b else off + (max(n, 1) - 1) * abs(step) + 1 + tail if n > 0 else off + tail
        base_shape[int(perm[ax])] = base_len
    inv = np.argsort(perm)

    def fn(base):
        v = base.transpose(*perm) if nd else base      # axis `ax` of v is base axis perm[ax]
        idx = []
        for ax in range(nd):
            n = shape[ax]
            if bcast[ax]:
                idx.append(slice(None))
                continue
            step, off = steps[ax], offs[ax]
            lo, hi = off, off + (n - 1) * abs(step) + 1 if n > 0 else off
            idx.append(slice(lo, hi, abs(step)))
        v = v[tuple(idx)]
        for ax in range(nd):
            if not bcast[ax] and steps[ax] < 0:
                v = np.flip(v, axis=ax)               # ndarray: s![lo..hi;-k] = the range, reversed
        if any(bcast):
            v = np.broadcast_to(v, shape)
        assert v.shape == tuple(shape), (v.shape, shape)
        return v

    del inv
    return tuple(base_shape), fn


def random_shape(rng):
    nd = int(rng.integers(1, 6))
    big = int(rng.integers(0, nd))
    shape = [int(rng.integers(1, 7)) for _ in range(nd)]
    shape[big] = int(rng.choice([1, 7, 33, 64, 130, 257]))
    if nd >= 2 and rng.random() < 0.5:
        shape[int(rng.integers(0, nd))] = int(rng.choice([32, 48, 96]))
    if rng.random() < 0.05:
        shape[int(rng.integers(0, nd))] = 0
    return shape


# NDB_FUZZ_SEEDS=N widens the sweep (used for long offline runs; the default keeps the suite short)
@pytest.mark.parametrize("seed", range(int(os.environ.get("NDB_FUZZ_SEEDS", 24))))
def test_fuzz_zip_views(nd, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        shape = random_shape(rng)
        dtype = [np.float32, np.float64, np.int32, np.int64][int(rng.integers(0, 4))]
        opname, op, nin = [("ADD", 2, 2), ("MUL", 4, 2), ("SUB", 3, 2), ("COPY", 0, 1), ("MUL_ADD", 7, 3), ("NEG", 15, 1)][int(rng.integers(0, 6))]
        ob, ofn = random_view_fn(rng, shape, allow_broadcast=False)
        mo = Mirror(nd, rand(rng, ob, dtype))
        ho, do = mo.view(ofn)
        hins, dins = [], []
        for _i in range(nin):
            ib, ifn = random_view_fn(rng, shape, allow_broadcast=True)
            m = Mirror(nd, rand(rng, ib, dtype))
            h, d = m.view(ifn)
            hins.append(h)
            dins.append(d)
        assert oracle.zip_(op, ho, hins, []) == 0
        from ndarray_b200.array import _zip_call
        _zip_call(do.ctx, op, do, dins, [])
        assert_bit_equal(mo.download(), mo.host, f"seed {seed} {opname} {np.dtype(dtype).name} shape {shape} out strides {ho.strides} "
                                                 f"in strides {[h.strides for h in hins]}")


@pytest.mark.parametrize("seed", range(int(os.environ.get("NDB_FUZZ_SEEDS", 24)) // 2))
def test_fuzz_reduce_axis_views(nd, oracle, seed):
    from ndarray_b200 import Axis
    rng = np.random.default_rng(5000 + seed)
    for _ in range(10):
        shape = random_shape(rng)
        shape = [max(s, 1) for s in shape]
        dtype = [np.float32, np.float64, np.int64][int(rng.integers(0, 3))]
        ib, ifn = random_view_fn(rng, shape, allow_broadcast=True)
        host = rand(rng, ib, dtype)
        if np.dtype(dtype).kind == "i":
            host = (host >> 8).astype(dtype)
        m = Mirror(nd, host)
        hv, dv = m.view(ifn)
        axis = int(rng.integers(0, len(shape)))
        got = dv.sum_axis(Axis(axis)).to_numpy()
        st, want = oracle.reduce_axis(0, hv, axis)
        assert st == 0
        what = f"seed {seed} sum_axis({axis}) {np.dtype(dtype).name} shape {shape} strides {hv.strides}"
        if np.dtype(dtype).kind == "i":
            assert np.array_equal(got, want), what
        else:
            tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(shape[axis]) * max(1.0, float(np.abs(hv).max())) * 4
            assert np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)), initial=0.0) <= tol * max(1, shape[axis] ** 0.5), what
            if shape[axis] >= 1:
                gv = dv.var_axis(Axis(axis), 0.0).to_numpy()
                st, wv = oracle.reduce_axis(5, hv, axis, 0.0)
                assert st == 0
                assert np.max(np.abs(gv.astype(np.float64) - wv.astype(np.float64)), initial=0.0) <= (1e-4 if dtype == np.float32 else 1e-11), what


@pytest.mark.parametrize("seed", range(int(os.environ.get("NDB_FUZZ_SEEDS", 24)) // 2))
def test_fuzz_folds_over_views(nd, oracle, seed):
    """Zip::fold forms on random views: fold of op(inputs) (ndb_zip_fold), fold of the written values (ndb_zip_fold_write:
    fused when dense, Zip + reduction otherwise) and fold_while over the traversal order - integers exact, the written arrays
    bit-exact, float sums to the north-star tolerance."""
    from ndarray_b200 import Zip, Op
    rng = np.random.default_rng(9000 + seed)
    for _ in range(8):
        shape = [max(s, 1) for s in random_shape(rng)]
        dtype = [np.float32, np.float64, np.int32, np.int64][int(rng.integers(0, 4))]
        kind = np.dtype(dtype).kind
        ob, ofn = random_view_fn(rng, shape, allow_broadcast=False)
        mo = Mirror(nd, rand(rng, ob, dtype) if kind == "f" else (rand(rng, ob, dtype) >> 12).astype(dtype))
        ho, do = mo.view(ofn)
        hins, dins = [], []
        for _i in range(2):
            ib, ifn = random_view_fn(rng, shape, allow_broadcast=True)
            base = rand(rng, ib, dtype) if kind == "f" else (rand(rng, ib, dtype) >> 12).astype(dtype)
            m = Mirror(nd, base)
            h, d = m.view(ifn)
            hins.append(h)
            dins.append(d)
        n = int(np.prod(shape))
        tol = (1e-5 if dtype == np.float32 else 1e-12) * np.sqrt(n) * 2.0
        # fold of op(inputs): nothing is written
        want_vals = hins[0] + hins[1]
        got = Zip.from_(dins[0]).and_(dins[1]).fold(None, Op.ADD, "sum")
        if kind == "i":
            with np.errstate(over="ignore"):
                assert got == want_vals.sum(dtype=dtype), (seed, shape)
        else:
            assert abs(float(got) - want_vals.astype(np.float64).sum()) <= tol * max(1.0, float(np.abs(want_vals).mean()))
        assert Zip.from_(dins[0]).and_(dins[1]).fold(None, Op.ADD, "max") == want_vals.max()
        # fold of the written values
        assert oracle.zip_(4, ho, hins, []) == 0
        got = Zip.from_(do).and_(dins[0]).and_(dins[1]).for_each_fold("a=b*c", "sum")
        assert_bit_equal(mo.download(), mo.host, f"seed {seed} fold_write {np.dtype(dtype).name} shape {shape}")
        if kind == "i":
            with np.errstate(over="ignore"):
                assert got == ho.sum(dtype=dtype)
        else:
            assert abs(float(got) - ho.astype(np.float64).sum()) <= tol * max(1.0, float(np.abs(ho).mean()))
        # fold_while over the traversal order of a C-order copy: stop at the first value below the threshold
        flat = np.ascontiguousarray(hins[0])
        from ndarray_b200 import Array
        dflat = Array.from_numpy(flat)
        thr = dtype(np.sort(flat.ravel())[max(0, flat.size // 3 - 1)]) if flat.size else dtype(0)
        seq = flat.ravel()
        hit = np.nonzero(seq < thr)[0]
        acc, done = Zip.from_(dflat).fold_while(dtype(0), Op.COPY, "max", 27, thr)
        if len(hit):
            first = int(hit[0])
            assert done and acc == (max(dtype(0), seq[:first].max()) if first else dtype(0)), (seed, shape)
        else:
            assert not done and acc == max(dtype(0), seq.max())
