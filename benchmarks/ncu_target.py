"""One small workload per invocation, for `ncu --set full` captures (kept short: ncu replays every kernel ~40x).

  python benchmarks/ncu_target.py zip|c5|combined|reduce|gemm64|gemm32|gemm64sk
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array, Axis, Zip, s_  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "zip"
    ctx = nd.default_context()
    rng = np.random.default_rng(0)
    if which == "zip":
        n = 512
        x = Array.from_numpy(rng.random((n, n, n), dtype=np.float32))
        y = Array.from_numpy(rng.random((n, n, n), dtype=np.float32))
        row = Array.from_numpy(rng.random(n, dtype=np.float32))
        out = Array.uninit((n, n, n), np.float32)
        for _ in range(3):
            Zip.from_(out).and_(x.t()).and_(y).for_each("a=b+c")
            Zip.from_(out.slice(s_[: n // 2, :, :])).and_(x.slice(s_[::-2, :, :])).and_(y.slice(s_[::2, :, :])).for_each("a=b+c")
            Zip.from_(out).and_(x).and_broadcast(row).for_each("a=b*c")
            Zip.from_(out).and_(x.slice(s_[:, :, ::-1])).and_(y).for_each("a=b+c")
    elif which == "c5":
        # one 8-GPU shard of the bench workload (config 5): a = b*c + a over 2^29 f32, then the global sum
        import ctypes as C
        rows, cols = 8192, 65536
        a, b, c = (Array.uninit((rows, cols), np.float32) for _ in range(3))
        for arr, seed in ((a, 1), (b, 2), (c, 3)):
            nd.lib().ndb_fill_counter_rng(ctx.handle, C.byref(arr.v), seed, 0)
        for _ in range(2):
            Zip.from_(a).and_(b).and_(c).for_each("a=b*c+a")
            a.sum()
    elif which == "combined":
        n = 512
        x = Array.from_numpy(rng.random((n, n, n), dtype=np.float32))
        y = Array.from_numpy(rng.random((n, n, n), dtype=np.float32))
        row = Array.from_numpy(rng.random(n, dtype=np.float32))
        out = Array.uninit((n, n, n), np.float32)
        o_half = out.slice(s_[: n // 2, :, :])
        rowb = row.broadcast((n // 2, n, n))
        xts, ys = x.t().slice(s_[::-2, :, :]), y.slice(s_[::2, :, :])
        for _ in range(3):
            Zip.from_(o_half).and_(ys).and_(rowb).and_(xts).for_each("a=b*c+d")
    elif which == "reduce":
        n = 16384
        f = Array.from_numpy(rng.standard_normal((n, n)))
        for _ in range(2):
            for ax in (0, 1):
                f.sum_axis(Axis(ax))
                f.var_axis(Axis(ax), 1.0)
            f.sum()
        a = Array.from_numpy(rng.random((4096, 4096), dtype=np.float32))
        b = Array.from_numpy(rng.random((4096, 4096), dtype=np.float32))
        c = Array.from_numpy(rng.random((4096, 4096), dtype=np.float32))
        for _ in range(2):
            Zip.from_(a).and_(b).and_(c).for_each("a=b*c+a")
            a.sum_axis(Axis(0))
            a.sum_axis(Axis(1))
    elif which in ("gemm64", "gemm32"):
        dt = np.float64 if which == "gemm64" else np.float32
        n = 4096
        A = Array.from_numpy(rng.standard_normal((n, n)).astype(dt))
        B = Array.from_numpy(rng.standard_normal((n, n)).astype(dt))
        Cm = Array.zeros((n, n), dt)
        for _ in range(2):
            nd.general_mat_mul(1.0, A, B, 0.0, Cm)
    elif which == "gemm64sk":
        # a ragged last wave (256 tiles on 148 SMs): whole tiles, then stream-K
        import os
        n = 2048
        A = Array.from_numpy(rng.standard_normal((n, n)))
        B = Array.from_numpy(rng.standard_normal((n, n)))
        Cm = Array.zeros((n, n), np.float64)
        for mode in ("0", "1"):
            os.environ["NDB_GEMM_F64_STREAMK"] = mode
            for _ in range(2):
                nd.general_mat_mul(1.0, A, B, 0.0, Cm)
    ctx.sync()
    print("done", which, "launches", nd.lib().ndb_launch_count())


if __name__ == "__main__":
    main()
