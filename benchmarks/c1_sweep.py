"""C1 sum_axis (f32 4096^2, cold L2) under the reduction split knobs."""
import os
import statistics
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array, Axis  # noqa: E402

ctx = nd.default_context()
dev = torch.device("cuda", 0)
n = 4096
bufs = [torch.rand(n, n, device=dev) for _ in range(8)]  # 8 x 64 MiB rotate through the 126 MB L2
arrs = [Array.borrow(t.data_ptr(), (n, n), np.float32, "C", ctx, owner=t) for t in bufs]
outs = [Array.uninit((n,), np.float32, ctx=ctx) for _ in bufs]


def run_batch(ax, rounds=10):
    """8 back-to-back launches (one per 64 MiB buffer, so each input is cold in L2) between one event pair."""
    from ndarray_b200.array import Red
    import ctypes as C
    L = nd.lib()
    evs = [(ctx.event(), ctx.event()) for _ in range(rounds)]
    for a, b in evs:
        ctx.record(a)
        for i in range(8):
            L.ndb_reduce_axis(ctx.handle, Red.SUM, ax, C.byref(arrs[i].v), C.byref(outs[i].v), 0.0)
        ctx.record(b)
    ctx.sync()
    ms = [ctx.elapsed_ms(a, b) / 8 for a, b in evs]
    return statistics.median(ms) * 1e3, min(ms) * 1e3


def run(ax, iters=40):
    from ndarray_b200.array import Red
    import ctypes as C
    L = nd.lib()
    evs = [(ctx.event(), ctx.event()) for _ in range(iters)]
    for i in range(8):
        L.ndb_reduce_axis(ctx.handle, Red.SUM, ax, C.byref(arrs[i].v), C.byref(outs[i].v), 0.0)
    ctx.sync()
    for i, (a, b) in enumerate(evs):
        ctx.record(a)
        L.ndb_reduce_axis(ctx.handle, Red.SUM, ax, C.byref(arrs[i % 8].v), C.byref(outs[i % 8].v), 0.0)
        ctx.record(b)
    ctx.sync()
    ms = [ctx.elapsed_ms(a, b) for a, b in evs]
    return statistics.median(ms) * 1e3, min(ms) * 1e3


for wpb in ("8", "4", "2", "1"):
    os.environ["NDB_RED_WPB"] = wpb
    med, best = run(1)
    print(f"axis1 warps/block={wpb}: median {med:5.1f} us best {best:5.1f} us", flush=True)
os.environ.pop("NDB_RED_WPB")
med, best = run(1)
print(f"axis1 default: median {med:5.1f} us best {best:5.1f} us", flush=True)
med, best = run(0)
print(f"axis0 default: median {med:5.1f} us best {best:5.1f} us", flush=True)
for ax in (0, 1):
    med, best = run_batch(ax)
    print(f"axis{ax} back-to-back x8: median {med:5.2f} us per launch ({4 * n * n / med / 1e3:6.0f} GB/s), best {best:5.2f}", flush=True)
for wpb in ("8", "4", "2"):
    os.environ["NDB_RED_WPB"] = wpb
    med, best = run_batch(1)
    print(f"axis1 back-to-back warps/block={wpb}: {med:5.2f} us", flush=True)
os.environ.pop("NDB_RED_WPB")
for per_sm in ("2", "3", "4", "6", "8"):
    os.environ["NDB_RED_COLS_PER_SM"] = per_sm
    med, best = run_batch(0)
    print(f"axis0 back-to-back per_sm={per_sm}: {med:5.2f} us", flush=True)
os.environ.pop("NDB_RED_COLS_PER_SM")
for splits in ("4", "6", "8", "9", "12", "18", "37"):
    os.environ["NDB_RED_COLS_SPLITS"] = splits
    med, best = run_batch(0)
    print(f"axis0 back-to-back splits={splits}: {med:5.2f} us", flush=True)
os.environ.pop("NDB_RED_COLS_SPLITS")
os.environ["NDB_RED_TWO_KERNELS"] = "1"
med, best = run_batch(0)
print(f"axis0 back-to-back two kernels (no ticket): {med:5.2f} us", flush=True)
os.environ.pop("NDB_RED_TWO_KERNELS")
for wmax in ("2048", "8192"):
    os.environ["NDB_RED_WARP_MAX"] = wmax
    med, best = run_batch(1)
    print(f"axis1 back-to-back warp_max={wmax} ({'block' if wmax == '2048' else 'warp'} per row): {med:5.2f} us", flush=True)
os.environ.pop("NDB_RED_WARP_MAX")
for nb in ("296", "444", "592", "888"):
    os.environ["NDB_RED_LANES_BLOCKS"] = nb
    med, best = run_batch(1)
    print(f"axis1 back-to-back blocks={nb}: {med:5.2f} us", flush=True)
