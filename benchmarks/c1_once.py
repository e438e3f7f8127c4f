"""One cold sum_axis per axis on f32 4096^2 (ncu target: per-kernel durations without launch gaps)."""
import sys
from pathlib import Path
import ctypes as C
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array  # noqa: E402
from ndarray_b200.array import Red  # noqa: E402
ctx = nd.default_context()
n = 4096
bufs = [torch.rand(n, n, device="cuda") for _ in range(4)]
arrs = [Array.borrow(t.data_ptr(), (n, n), np.float32, "C", ctx, owner=t) for t in bufs]
out = Array.uninit((n,), np.float32, ctx=ctx)
for rep in range(3):
    for ax in (0, 1):
        for i in range(4):
            nd.lib().ndb_reduce_axis(ctx.handle, Red.SUM, ax, C.byref(arrs[i].v), C.byref(out.v), 0.0)
ctx.sync()
