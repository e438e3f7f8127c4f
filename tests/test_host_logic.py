"""CPU-only checks of the C-ABI library's host side: it loads, exports every symbol the header declares,
its view arithmetic agrees with the oracle (and therefore with ndarray's golden vectors), the Zip planner
picks the expected kernel classes, and compute entry points fail loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import ndarray_b200 as nd
from ndarray_b200 import _ffi
from ndarray_b200._ffi import View, ZipPlan
from oracle import oracle as orc

ROOT = Path(__file__).resolve().parent.parent
DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3}


def mk(ptr, dtype, dim, stride) -> View:
    v = View()
    v.ptr = ptr
    v.ndim = len(dim)
    v.dtype = DT[np.dtype(dtype)]
    for i, (d, s) in enumerate(zip(dim, stride)):
        v.dim[i] = d
        v.stride[i] = s
    return v


def c_strides(shape):
    st, c = [], 1
    for d in reversed(shape):
        st.append(c)
        c *= d
    return list(reversed(st))


def f_strides(shape):
    st, c = [], 1
    for d in shape:
        st.append(c)
        c *= d
    return st


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "ndb200.h").read_text()
    names = set(re.findall(r"\b((?:ndb|cblas)_[a-z0-9_]+)\s*\(", hdr))
    names -= {"ndb_zip_plan_t"}
    assert len(names) > 40
    lib = C.CDLL(str(_ffi.LIB_PATH))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in nd.lib().ndb_version()


def test_struct_layout_matches_header():
    assert C.sizeof(View) == 8 + 4 + 4 + 8 * 8 + 8 * 8
    assert C.sizeof(ZipPlan) == 4 + 4 + 64 + 9 * 64 + 72 + 4 + 4


def test_compute_fails_loudly_without_gpu():
    n = C.c_int()
    st = nd.lib().ndb_device_count(C.byref(n))
    if st == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nd.NdbError):
        nd.Context(0)
    a = mk(4096, np.float32, (4,), (1,))
    parts = (View * 2)(a, a)
    assert nd.lib().ndb_zip(None, 0, None, 2, parts) != 0  # null context: no fallback
    # every other compute entry point refuses a null context the same way (there is no CPU path to fall into)
    L = nd.lib()
    out = np.zeros(4, np.float32)
    hp = C.c_void_p(out.ctypes.data)
    assert L.ndb_reduce_all(None, 0, C.byref(a), hp) != 0
    assert L.ndb_reduce_all_dev(None, 0, C.byref(a), hp) != 0
    assert L.ndb_reduce_all_peer(None, 0, C.byref(a), hp) != 0
    assert L.ndb_reduce_axis(None, 0, 0, C.byref(a), C.byref(a), 0.0) != 0
    assert L.ndb_reduce_axis_peer(None, 0, 0, C.byref(a), C.byref(a), 0.0, 4) != 0
    assert L.ndb_accumulate_axis(None, 2, 0, C.byref(a)) != 0
    assert L.ndb_cast(None, C.byref(a), C.byref(a)) != 0
    assert L.ndb_dot(None, C.byref(a), C.byref(a), hp) != 0
    assert L.ndb_peer_allreduce(None, 0, 0, hp, hp, 4) != 0
    assert L.ndb_peer_allgather(None, 0, hp, hp, 4) != 0
    assert L.ndb_peer_window_create(None, 0, 1, 1 << 10, hp) != 0
    assert L.ndb_peer_window_attach(None, hp) != 0


@pytest.mark.parametrize("spec", [(1, 3, -1), (1, None, -2), (0, 4, -2), (0, None, 2), (-3, -1, 1), (2, 2, 1), (0, 0, -1),
                                  (4, None, 1), (1, None, 3), (0, None, -3)])
def test_slice_axis_matches_oracle(spec):
    s, e, st = spec
    for length, stride in [(4, 1), (7, -3), (1, 5), (10, 2)]:
        v = mk(10_000, np.float64, (3, length), (100, stride))
        o = orc.raw_view(10_000, np.float64, (3, length), (100, stride))
        r1 = nd.lib().ndb_view_slice_axis(C.byref(v), 1, s, 0 if e is None else e, 0 if e is None else 1, st)
        r2 = orc.slice_axis(o, 1, s, e, st)
        assert (r1 == 0) == (r2 == 0)
        if r1 == 0:
            assert (v.ptr, v.dim[1], v.stride[1]) == (o.ptr, o.dim[1], o.stride[1])


def test_slice_errors():
    v = mk(0, np.float32, (4,), (1,))
    assert nd.lib().ndb_view_slice_axis(C.byref(v), 0, 5, 0, 0, 1) == _ffi.ERR_OOB
    assert b"Slice begin 5 is past end of axis of length 4" in nd.lib().ndb_last_error()
    assert nd.lib().ndb_view_slice_axis(C.byref(v), 0, 0, 0, 0, 0) == _ffi.ERR_INVALID
    assert nd.lib().ndb_view_slice_axis(C.byref(v), 1, 0, 0, 0, 1) == _ffi.ERR_OOB


def test_broadcast_matches_reference_rules():  # tests/broadcast.rs:4-33
    def bc(from_shape, to_shape):
        v = mk(64, np.int64, from_shape, c_strides(from_shape))
        out = View()
        arr = (C.c_int64 * max(len(to_shape), 1))(*to_shape)
        st = nd.lib().ndb_view_broadcast(C.byref(v), len(to_shape), arr, C.byref(out))
        return st, out

    assert bc((2, 1, 2, 1), (2, 4, 2, 2))[0] == 0
    assert bc((2, 1), (1,))[0] == _ffi.ERR_SHAPE
    assert bc((2, 1), ())[0] == _ffi.ERR_SHAPE
    assert bc((2, 1), (32, 1, 2))[0] == _ffi.ERR_SHAPE
    st, out = bc((2, 1), (32, 2, 5))
    assert st == 0 and [out.stride[i] for i in range(3)] == [0, 1, 0]
    st, out = bc((), (7, 2, 9))
    assert st == 0 and [out.stride[i] for i in range(3)] == [0, 0, 0]
    st, out = bc((512,), (512, 512, 512))
    assert st == 0 and [out.stride[i] for i in range(3)] == [0, 0, 1]


def test_layout_contiguity_min_stride_match_oracle():
    rng = np.random.default_rng(0)
    base = np.zeros((4, 5, 6), np.float32)
    views = [base, base.T, base.transpose(1, 0, 2), base[::-1], base[:, ::2], base[:, :, ::-1], base[1:, :, :],
             base[:, :1, :], base[:, :, :1], base[:1], np.zeros((5, 5), np.float32, order="F")[1:],
             np.zeros(7, np.float32)[::-1], np.zeros((3, 3), np.float32)[:, 1], np.broadcast_to(np.zeros((3, 3), np.float32), (8, 3, 3))]
    for a in views:
        o = orc.view(a)
        v = mk(o.ptr, a.dtype, a.shape, [s // a.itemsize for s in a.strides])
        assert nd.lib().ndb_view_layout(C.byref(v)) == orc.layout(a), a.strides
        assert bool(nd.lib().ndb_view_is_contiguous(C.byref(v))) == orc.is_contiguous(a), a.strides
        assert nd.lib().ndb_view_min_stride_axis(C.byref(v)) == orc.min_stride_axis(a), a.strides


def test_invert_permute_split():
    v = mk(1000, np.float32, (3, 4), (4, 1))
    assert nd.lib().ndb_view_invert_axis(C.byref(v), 0) == 0
    assert (v.ptr, v.stride[0]) == (1000 + 2 * 4 * 4, -4)
    out = View()
    perm = (C.c_int * 2)(1, 0)
    assert nd.lib().ndb_view_permute(C.byref(v), perm, C.byref(out)) == 0
    assert (out.dim[0], out.dim[1], out.stride[0], out.stride[1]) == (4, 3, 1, -4)
    l, r = View(), View()
    v2 = mk(1000, np.float64, (10, 4), (4, 1))
    assert nd.lib().ndb_view_split_at(C.byref(v2), 0, 3, C.byref(l), C.byref(r)) == 0
    assert (l.dim[0], r.dim[0], r.ptr) == (3, 7, 1000 + 3 * 4 * 8)
    assert nd.lib().ndb_view_split_at(C.byref(v2), 0, 10, C.byref(l), C.byref(r)) == 0
    assert (l.dim[0], r.dim[0], r.ptr) == (10, 0, 1000)  # never points past the allocation
    assert nd.lib().ndb_view_split_at(C.byref(v2), 0, 11, C.byref(l), C.byref(r)) == _ffi.ERR_OOB


def plan(op, *views):
    parts = (View * len(views))(*views)
    pl = ZipPlan()
    st = nd.lib().ndb_zip_plan(op, len(views), parts, C.byref(pl))
    return st, pl


def test_planner_kernel_classes():
    A = 1 << 20  # 16-byte aligned fake addresses
    n = 512
    c3 = mk(A, np.float32, (n, n, n), c_strides((n, n, n)))
    # config 1: three C-order 4096^2 arrays collapse to one flat run
    c2 = mk(A, np.float32, (4096, 4096), (4096, 1))
    st, pl = plan(nd.Op.MUL_ADD, c2, c2, c2, c2)
    assert st == 0 and pl.kernel == 1 and pl.ndim == 1 and pl.dim[0] == 4096 * 4096
    # unaligned base -> scalar flat
    st, pl = plan(nd.Op.ADD, mk(A + 4, np.float32, (64,), (1,)), mk(A, np.float32, (64,), (1,)), mk(A, np.float32, (64,), (1,)))
    assert st == 0 and pl.kernel == 2
    # config 2 (i): x.t() + y -> tile transpose on the axis where x is dense
    xt = mk(A, np.float32, (n, n, n), f_strides((n, n, n)))
    st, pl = plan(nd.Op.ADD, c3, xt, c3)
    assert st == 0 and pl.kernel == 5 and pl.ndim == 3 and pl.tile_axis == 0 and pl.staged_mask == 0b010
    # config 2 (ii): s![..;-2, .., ..] / s![..;2, .., ..] -> two collapsed axes, vector rows
    xs = mk(A + 511 * n * n * 4, np.float32, (256, n, n), (-2 * n * n, n, 1))
    ys = mk(A, np.float32, (256, n, n), (2 * n * n, n, 1))
    out = mk(A, np.float32, (256, n, n), c_strides((256, n, n)))
    st, pl = plan(nd.Op.ADD, out, xs, ys)
    assert st == 0 and pl.kernel == 3 and pl.ndim == 2 and list(pl.dim[:2]) == [256, n * n]
    assert list(pl.stride[1][:2]) == [-2 * n * n, 1]
    # config 2 (iii): broadcast row (0,0,1) -> rows collapse to [n*n, n]
    row = mk(A, np.float32, (n, n, n), (0, 0, 1))
    st, pl = plan(nd.Op.MUL, c3, c3, row)
    assert st == 0 and pl.kernel == 3 and pl.ndim == 2 and list(pl.dim[:2]) == [n * n, n]
    assert list(pl.stride[2][:2]) == [0, 1]
    # inner-axis reversal: s![.., .., ..;-1] input stays on the 128-bit path (lanes reversed in registers) when the
    # mirrored rows are aligned, else the scalar rows kernel
    xr = mk(A + (n - 1) * 4, np.float32, (n, n, n), (n * n, n, -1))
    st, pl = plan(nd.Op.ADD, c3, xr, c3)
    assert st == 0 and pl.kernel == 3 and pl.stride[1][pl.ndim - 1] == -1
    xr2 = mk(A + (n - 2) * 4, np.float32, (n, n, n), (n * n, n, -1))
    st, pl = plan(nd.Op.ADD, c3, xr2, c3)
    assert st == 0 and pl.kernel == 4 and pl.stride[1][pl.ndim - 1] == -1
    # a backwards OUTPUT is flipped for every producer, so it becomes a flat forward run
    o_rev = mk(A + (4096 - 16) * 4 + 12, np.float32, (4096,), (-1,))
    i_rev = mk(A + (4096 - 16) * 4 + 12, np.float32, (4096,), (-1,))
    st, pl = plan(nd.Op.COPY, o_rev, i_rev)
    assert st == 0 and pl.kernel in (1, 2) and pl.stride[0][0] == 1 and pl.offset[0] == -4095
    # empty
    st, pl = plan(nd.Op.COPY, mk(A, np.float32, (3, 0), (0, 0)), mk(A, np.float32, (3, 0), (0, 0)))
    assert st == 0 and pl.kernel == 0


def test_planner_errors():
    A = 1 << 20
    a = mk(A, np.float32, (5, 7), (7, 1))
    b = mk(A, np.float32, (6, 7), (7, 1))
    st, _ = plan(nd.Op.COPY, a, b)
    assert st == _ffi.ERR_SHAPE
    assert b"Zip: Producer dimension mismatch, expected: [5, 7], got: [6, 7]" in nd.lib().ndb_last_error()
    st, _ = plan(nd.Op.ADD, a, a)  # arity
    assert st == _ffi.ERR_INVALID
    st, _ = plan(nd.Op.COPY, mk(A, np.float32, (5, 7), (0, 1)), a)  # output aliases itself
    assert st == _ffi.ERR_ALIASING
    st, _ = plan(nd.Op.COPY, a, mk(A, np.float64, (5, 7), (7, 1)))
    assert st == _ffi.ERR_UNSUPPORTED


def test_prefer_f_matches_oracle():
    b = np.zeros((5, 10), np.float32)
    cases = [[b, b], [b.T, b.T], [b, b.T], [b[:, ::2], b.T[::2].T], [np.zeros((5, 5), np.float32, order="F")[1:], b[:4, :5].T.T]]
    for ins in cases:
        views = (View * len(ins))(*[mk(orc.view(x).ptr, x.dtype, x.shape, [s // x.itemsize for s in x.strides]) for x in ins])
        assert bool(nd.lib().ndb_zip_prefer_f(len(ins), views)) == orc.zip_prefer_f(ins)


def test_merge_axes_golden():  # tests/array.rs:780-868
    L = nd.lib()

    def sliced(shape, strides, specs):
        v = mk(1 << 20, np.int32, shape, strides)
        for ax, sp in enumerate(specs):
            if sp is not None:
                s, e, st = sp
                assert L.ndb_view_slice_axis(C.byref(v), ax, s, 0 if e is None else e, 0 if e is None else 1, st) == 0
        return v

    def check(shape, strides, specs, take, into, expect):
        v = sliced(shape, strides, specs)
        before = ([v.dim[i] for i in range(4)], [v.stride[i] for i in range(4)])
        merged_len = v.dim[take] * v.dim[into]
        r = L.ndb_view_merge_axes(C.byref(v), take, into)
        assert bool(r) == expect, (specs, take, into)
        if expect:
            assert v.dim[take] == (0 if merged_len == 0 else 1) and v.dim[into] == merged_len
        else:
            assert ([v.dim[i] for i in range(4)], [v.stride[i] for i in range(4)]) == before

    shp = (3, 4, 5, 4)
    cs = c_strides(shp)
    full = [None] * 4
    merged_full = {(0, 1), (1, 2), (2, 3)}
    for t in range(4):
        for i in range(4):
            check(shp, cs, full, t, i, (t, i) in merged_full)
    last2 = [None, None, None, (0, None, 2)]
    for (t, i, e) in [(0, 1, True), (1, 0, False), (1, 2, True), (2, 1, False), (2, 3, True), (3, 2, False)]:
        check(shp, cs, last2, t, i, e)
    last3 = [None, None, None, (0, 3, 1)]
    for (t, i, e) in [(0, 1, True), (1, 0, False), (1, 2, True), (2, 1, False), (2, 3, False)]:
        check(shp, cs, last3, t, i, e)
    ax2 = [None, None, (0, None, 2), None]
    for (t, i, e) in [(0, 1, True), (1, 0, False), (1, 2, False), (2, 3, False)]:
        check(shp, cs, ax2, t, i, e)
    ax1 = [None, (0, None, 2), None, None]
    for (t, i, e) in [(0, 1, True), (1, 0, False), (1, 2, False), (2, 3, True), (3, 2, False)]:
        check(shp, cs, ax1, t, i, e)
    shp = (3, 1, 5, 1)
    fs = f_strides(shp)
    for (t, i) in [(0, 1), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3), (2, 1), (2, 3), (3, 0), (3, 1), (3, 2), (3, 3)]:
        check(shp, fs, ax2, t, i, True)
    shp = (3, 0, 5, 1)
    for (t, i) in [(0, 1), (1, 1), (2, 1), (3, 1), (1, 0), (1, 2), (1, 3)]:
        check(shp, [0, 0, 0, 0], ax2, t, i, True)


def test_collapse_axis():
    v = mk(1000, np.float64, (3, 4), (4, 1))
    assert nd.lib().ndb_view_collapse_axis(C.byref(v), 0, 2) == 0
    assert (v.ptr, v.dim[0], v.dim[1]) == (1000 + 2 * 4 * 8, 1, 4)
    assert nd.lib().ndb_view_collapse_axis(C.byref(v), 1, 4) == _ffi.ERR_OOB


def test_planner_combined_views_and_config5():
    """Config 2 (iv): all three views in one Zip -> the tile kernel with a stride -2 staged producer; config 5: 2^32
    elements stay one flat run with 64-bit extents; nine producers (tests/azip.rs:481) plan as one flat Zip."""
    A = 1 << 20
    n = 512
    out = mk(A, np.float32, (n // 2, n, n), c_strides((n // 2, n, n)))
    ys = mk(A, np.float32, (n // 2, n, n), (2 * n * n, n, 1))
    rowb = mk(A, np.float32, (n // 2, n, n), (0, 0, 1))
    # x.t() has strides (1, n, n*n); [::-2] on axis 0 -> stride -2 from the last element of the axis
    xts = mk(A + (n - 1) * 4, np.float32, (n // 2, n, n), (-2, n, n * n))
    st, pl = plan(nd.Op.MUL_ADD, out, ys, rowb, xts)
    assert st == 0 and pl.kernel == 5 and pl.staged_mask == 0b1000
    k = pl.tile_axis
    assert pl.stride[3][k] == -2 and pl.stride[3][pl.ndim - 1] == n * n
    big = mk(A, np.float32, (65536, 65536), (65536, 1))
    st, pl = plan(nd.Op.MUL_ADD, big, big, big, big)
    assert st == 0 and pl.kernel == 1 and pl.ndim == 1 and pl.dim[0] == 1 << 32
    v = mk(A, np.int64, (7, 11), (11, 1))
    st, pl = plan(17, *([v] * 9))  # ADD_N: out + 8 inputs
    assert st == 0 and pl.kernel in (1, 2) and pl.dim[0] == 77
    st, pl = plan(17, *([v] * 10))
    assert st != 0  # more than 9 producers: ndarray's Zip stops at 9 too


def test_rust_ffi_declares_exactly_the_header():
    """rust/ndarray-b200/src/ffi.rs cannot be compiled here (no rustc), so its agreement with include/ndb200.h is
    checked mechanically: the same set of symbols, each with the same number of parameters; and every `ffi::` call
    in lib.rs names a declared symbol."""
    hdr = (ROOT / "include" / "ndb200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b((?:ndb|cblas)_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    rs = (ROOT / "rust" / "ndarray-b200" / "src" / "ffi.rs").read_text()
    rust = {}
    for m in re.finditer(r"pub fn ((?:ndb|cblas)_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", rs):
        args = m.group(2).strip()
        rust[m.group(1)] = 0 if not args else args.count(",") + 1
    assert set(rust) == set(decl), (sorted(set(decl) - set(rust)), sorted(set(rust) - set(decl)))
    wrong = {n: (decl[n], rust[n]) for n in decl if decl[n] != rust[n]}
    assert not wrong, wrong
    lib_rs = (ROOT / "rust" / "ndarray-b200" / "src" / "lib.rs").read_text()
    used = set(re.findall(r"ffi::((?:ndb|cblas)_[a-z0-9_]+)\s*\(", lib_rs))
    assert used <= set(rust), sorted(used - set(rust))
    assert len(used) >= 40  # the crate actually covers the ABI, not a corner of it


def test_streamk_partition_invariants():
    """A model of the range arithmetic shared by `streamk_launch` (host: region size, grid) and `gemm_f64_tma_streamk`
    (device: ticket -> [it0, it1), segments, owner / contributor roles) in csrc/gemm_f64_tma.cu, checked over random shapes
    for the properties the hand-over protocol relies on: every (tile, k-slab) is computed exactly once; every tile has
    exactly one owner (the CTA holding its first slab); only a CTA's FIRST segment can start inside a tile, so a CTA parks at
    most one partial and never waits before doing so; the contributors an owner walks (tickets w+1.. while their range
    starts before the tile's end) are exactly the CTAs that parked a partial for that tile, each consumed once - so every
    flag raised is lowered again and the buffer is clean when the grid retires."""
    import random

    def check(out_tiles, kt, sms):
        sk_tiles = sms + out_tiles % sms if out_tiles > sms else out_tiles
        iters = sk_tiles * kt
        grid = sms + out_tiles - sk_tiles
        cover, parked, owners = {}, {}, {}
        for w in range(grid):
            if w < sms:
                it0, it1 = iters * w // sms, iters * (w + 1) // sms
            else:
                it0 = (sk_tiles + (w - sms)) * kt
                it1 = it0 + kt
            assert it1 > it0
            it, first = it0, True
            while it < it1:
                tile = it // kt
                kb = it - tile * kt
                nseg = min(it1 - it, kt - kb)
                for k in range(kb, kb + nseg):
                    assert (tile, k) not in cover
                    cover[(tile, k)] = w
                if kb == 0:
                    assert tile not in owners
                    owners[tile] = (w, nseg)
                else:
                    assert first and w not in parked
                    parked[w] = tile
                it += nseg
                first = False
        assert len(cover) == out_tiles * kt and len(owners) == out_tiles
        consumed = set()
        for tile, (w, nseg) in owners.items():
            if nseg < kt:
                c = w + 1
                while c < sms and iters * c // sms < (tile + 1) * kt:
                    assert parked.get(c) == tile and c not in consumed
                    consumed.add(c)
                    c += 1
        assert consumed == set(parked)

    rng = random.Random(1)
    for sms in (148, 132, 2, 3, 7):
        for _ in range(150):
            out_tiles, kt = rng.randint(1, 5 * sms + 3), rng.randint(1, 300)
            sk_tiles = sms + out_tiles % sms if out_tiles > sms else out_tiles
            if sk_tiles * kt >= 2 * sms:  # the launcher's floor: no CTA gets an empty range
                check(out_tiles, kt, sms)
    for out_tiles, kt in ((256, 64), (400, 80), (196, 128), (512, 256), (56, 17), (15, 64), (600, 8), (182, 9)):
        check(out_tiles, kt, 148)  # the shapes of benchmarks/gemm_streamk.py and test_f64_stream_k_matches_whole_tile_kernel
