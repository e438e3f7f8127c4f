"""C1 sum_axis (f32 4096^2, cold L2), back to back: knobs given on the command line as ENV=VAL pairs per run."""
import os
import statistics
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array  # noqa: E402
from ndarray_b200.array import Red  # noqa: E402
import ctypes as C  # noqa: E402

ctx = nd.default_context()
dev = torch.device("cuda", 0)
n = 4096
bufs = [torch.rand(n, n, device=dev) for _ in range(8)]
arrs = [Array.borrow(t.data_ptr(), (n, n), np.float32, "C", ctx, owner=t) for t in bufs]
outs = [Array.uninit((n,), np.float32, ctx=ctx) for _ in bufs]
L = nd.lib()


def run_batch(ax, rounds=15):
    evs = [(ctx.event(), ctx.event()) for _ in range(rounds)]
    for a, b in evs:
        ctx.record(a)
        for i in range(8):
            L.ndb_reduce_axis(ctx.handle, Red.SUM, ax, C.byref(arrs[i].v), C.byref(outs[i].v), 0.0)
        ctx.record(b)
    ctx.sync()
    ms = [ctx.elapsed_ms(a, b) / 8 for a, b in evs]
    return statistics.median(ms) * 1e3


for spec in sys.argv[1:]:
    ax, _, envs = spec.partition(":")
    kv = dict(e.split("=") for e in envs.split(",") if e)
    os.environ.update(kv)
    us = run_batch(int(ax))
    print(f"{os.environ.get('NDB200_LIB', 'default').split('_')[-1]:>12} axis{ax} {envs:40s} {us:6.2f} us  {4 * n * n / us / 1e3:6.0f} GB/s", flush=True)
    for k in kv:
        os.environ.pop(k)
