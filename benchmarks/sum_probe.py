"""Whole-array sum timing (C3 f64 16384^2 and f32), host-synchronous API vs device-resident result."""
import ctypes as C
import statistics
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array  # noqa: E402
ctx = nd.default_context()
L = nd.lib()
for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
    t = torch.randn(16384, 16384, device="cuda", dtype=tdt)
    a = Array.borrow(t.data_ptr(), (16384, 16384), dt, "C", ctx, owner=t)
    acc = Array.zeros((), dt, ctx=ctx)
    for name, fn in (("sum() to host", lambda: a.sum()), ("sum to device", lambda: L.ndb_reduce_all_dev(ctx.handle, 0, C.byref(a.v), C.c_void_p(acc.v.ptr)))):
        for _ in range(3):
            fn()
        ctx.sync()
        evs = [(ctx.event(), ctx.event()) for _ in range(20)]
        for e0, e1 in evs:
            ctx.record(e0)
            fn()
            ctx.record(e1)
        ctx.sync()
        ms = statistics.median(ctx.elapsed_ms(e0, e1) for e0, e1 in evs)
        print(f"{np.dtype(dt).name} {name:16s} {ms * 1e3:8.1f} us  {t.numel() * t.element_size() / ms / 1e6:7.0f} GB/s", flush=True)
