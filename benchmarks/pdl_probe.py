"""Does programmatic dependent launch help or hurt long kernels?  One 8-GPU shard of config 5 (2^29 f32 per array):
back-to-back sums, back-to-back Zips, and whole steps (Zip + sum), 20 in a row between one event pair."""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array, Zip  # noqa: E402
ctx = nd.default_context()
L = nd.lib()
rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 65536
a, b, c = (Array.uninit((rows, cols), np.float32, ctx=ctx) for _ in range(3))
for arr, seed in ((a, 1), (b, 2), (c, 3)):
    L.ndb_fill_counter_rng(ctx.handle, C.byref(arr.v), seed, 0)
acc = Array.zeros((), np.float32, ctx=ctx)
z = Zip.from_(a).and_(b).and_(c)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ctx.sync()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(n):
        fn()
    ctx.record(e1)
    ctx.sync()
    return ctx.elapsed_ms(e0, e1) / n


def s():
    L.ndb_reduce_all_dev(ctx.handle, 0, C.byref(a.v), C.c_void_p(acc.v.ptr))


def zf():
    z.for_each("a=b*c+a")


def step():
    zf()
    s()


print(f"rows {rows}: sum x100  {timed(s, 100):.4f} ms   zip x20  {timed(zf):.4f} ms   step x20  {timed(step):.4f} ms   sum x100 again {timed(s, 100):.4f} ms", flush=True)
