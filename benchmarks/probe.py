"""Quick on-box probes used while tuning (not part of the bench contract)."""
import os
import sys
import statistics
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array, Zip, s_, Op, Axis  # noqa: E402


def timeit(ctx, fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    evs = [(ctx.event(), ctx.event()) for _ in range(iters)]
    for a, b in evs:
        ctx.record(a)
        fn()
        ctx.record(b)
    ctx.sync()
    ms = [ctx.elapsed_ms(a, b) for a, b in evs]
    return statistics.median(ms), min(ms)


def main():
    ctx = nd.default_context()
    n = 512
    x = Array.zeros((n, n, n), np.float32)
    y = Array.zeros((n, n, n), np.float32)
    out = Array.uninit((n, n, n), np.float32)
    row = Array.zeros((n,), np.float32)
    cases = {
        "transposed x.t()+y": (lambda: Zip.from_(out).and_(x.t()).and_(y).for_each("a=b+c"), 3 * n ** 3 * 4),
        "perm (1,0,2)": (lambda: Zip.from_(out).and_(x.permuted_axes([1, 0, 2])).and_(y).for_each("a=b+c"), 3 * n ** 3 * 4),
        "perm (0,2,1)": (lambda: Zip.from_(out).and_(x.permuted_axes([0, 2, 1])).and_(y).for_each("a=b+c"), 3 * n ** 3 * 4),
        "bcast row": (lambda: Zip.from_(out).and_(x).and_broadcast(row).for_each("a=b*c"), 2 * n ** 3 * 4),
        "neg outer": (lambda: Zip.from_(out.slice(s_[:n // 2, :, :])).and_(x.slice(s_[::-2, :, :])).and_(y.slice(s_[::2, :, :])).for_each("a=b+c"), 3 * n ** 3 * 2),
        "neg inner": (lambda: Zip.from_(out).and_(x.slice(s_[:, :, ::-1])).and_(y).for_each("a=b+c"), 3 * n ** 3 * 4),
        "copy flat": (lambda: Zip.from_(out).and_(x).for_each("a=b"), 2 * n ** 3 * 4),
        "copy t": (lambda: Zip.from_(out).and_(x.t()).for_each("a=b"), 2 * n ** 3 * 4),
    }
    for env in ({}, {"NDB_TILE_YFAST": "1"}, {"NDB_TILE_SCALAR": "1"}):
        for k in ("NDB_TILE_YFAST", "NDB_TILE_SCALAR"):
            os.environ.pop(k, None)
        os.environ.update(env)
        print("env", env)
        for name, (fn, nbytes) in cases.items():
            if env and "t" not in name and "perm" not in name:
                continue
            med, best = timeit(ctx, fn)
            print(f"  {name:24s} {med*1e3:8.1f} us  {nbytes/med/1e6:8.1f} GB/s (best {nbytes/best/1e6:8.1f})")
    # C1-size reductions, cold L2 (four 64 MiB arrays in rotation), preallocated output
    import ctypes as C
    from ndarray_b200._ffi import lib, check
    arrs = [Array.zeros((4096, 4096), np.float32) for _ in range(4)]
    o = Array.zeros((4096,), np.float32)
    cnt = [0]
    for ax in (0, 1):
        def fn(ax=ax):
            cnt[0] += 1
            check(lib().ndb_reduce_axis(ctx.handle, 0, ax, C.byref(arrs[cnt[0] % 4].v), C.byref(o.v), 0.0))
        med, best = timeit(ctx, fn, 40, 4)
        print(f"  COLD sum_axis{ax} f32 4096^2 into: {med*1e3:8.1f} us {4096*4096*4/med/1e6:8.1f} GB/s")
    del arrs
    # reductions at C1 size with preallocated outputs
    a = Array.zeros((4096, 4096), np.float32)
    for ax in (0, 1):
        o = Array.zeros((4096,), np.float32)
        from ndarray_b200._ffi import lib, check
        import ctypes as C
        fn = lambda: check(lib().ndb_reduce_axis(ctx.handle, 0, ax, C.byref(a.v), C.byref(o.v), 0.0))
        med, best = timeit(ctx, fn, 50, 5)
        print(f"  sum_axis{ax} f32 4096^2 into: {med*1e3:8.1f} us {4096*4096*4/med/1e6:8.1f} GB/s")
        med, best = timeit(ctx, lambda: a.sum_axis(Axis(ax)), 50, 5)
        print(f"  sum_axis{ax} f32 4096^2 alloc: {med*1e3:8.1f} us {4096*4096*4/med/1e6:8.1f} GB/s")
    import ctypes as C
    from ndarray_b200._ffi import lib, check
    n = 16384
    for dt, ops in ((np.float64, (("sum", 0), ("var", 5), ("min", 2))), (np.int64, (("sum", 0), ("min", 2)))):
        big = Array.zeros((n, n), dt)
        o = Array.zeros((n,), dt)
        for name, red in ops:
            for ax in (0, 1):
                fn = lambda: check(lib().ndb_reduce_axis(ctx.handle, red, ax, C.byref(big.v), C.byref(o.v), 1.0))
                med, best = timeit(ctx, fn, 10, 2)
                print(f"  {name}_axis{ax} {np.dtype(dt).name} 16384^2 into: {med*1e3:8.1f} us {n*n*8/med/1e6:8.1f} GB/s")
        del big
    for dt in (np.float32, np.float64):
        nn = 1 << 28
        x1, y1 = Array.zeros((nn,), dt), Array.zeros((nn,), dt)
        med, best = timeit(ctx, lambda: x1.dot(y1), 10, 2)
        print(f"  dot {np.dtype(dt).name} 2^28: {med*1e3:8.1f} us {2*nn*np.dtype(dt).itemsize/med/1e6:8.1f} GB/s")
        med, best = timeit(ctx, lambda: x1.slice(s_[::2]).dot(y1.slice(s_[::-2])), 10, 2)
        print(f"  dot {np.dtype(dt).name} 2^27 strided: {med*1e3:8.1f} us {nn*np.dtype(dt).itemsize/med/1e6:8.1f} GB/s (algorithmic)")
        del x1, y1
    n = 16384
    for dt in (np.float32, np.float64):
        for order in ("C", "F"):
            A = Array.zeros((n, n), dt, order=order)
            x, y = Array.zeros((n,), dt), Array.zeros((n,), dt)
            med, best = timeit(ctx, lambda: nd.general_mat_vec_mul(1.0, A, x, 0.0, y), 10, 2)
            print(f"  gemv {np.dtype(dt).name} {n}^2 A={order}: {med*1e3:8.1f} us {n*n*np.dtype(dt).itemsize/med/1e6:8.1f} GB/s")
            del A
    if os.environ.get("NDB_PROBE_NO_GEMM"):
        return
    for dt in (np.float64, np.float32):
        for n in (2048, 4096, 8192):
            A, B, Cm = (Array.zeros((n, n), dt) for _ in range(3))
            med, best = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), 3, 1)
            print(f"  gemm {np.dtype(dt).name} {n}: {med:8.3f} ms {2*n**3/med/1e9:8.2f} TFLOP/s")
            if dt == np.float64:
                Af = Array.zeros((n, n), dt, order="F")
                med, best = timeit(ctx, lambda: nd.general_mat_mul(1.0, Af, B, 0.0, Cm), 3, 1)
                print(f"  gemm {np.dtype(dt).name} {n} A=F: {med:8.3f} ms {2*n**3/med/1e9:8.2f} TFLOP/s")


if __name__ == "__main__":
    main()
