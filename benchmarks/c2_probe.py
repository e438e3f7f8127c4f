"""Config 2 (iv): all three views in one Zip (stepped transposed producer -> the STEP2 tile kernel), plus the plain cases."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array, Zip, s_  # noqa: E402
from benchmarks.extras import _time  # noqa: E402
ctx = nd.default_context()
n = 512
tx, ty, trow = torch.rand(n, n, n, device="cuda") * 2 - 1, torch.rand(n, n, n, device="cuda") * 2 - 1, torch.rand(n, device="cuda")
x, y, row = (Array.borrow(t.data_ptr(), tuple(t.shape), np.float32, "C", ctx, owner=t) for t in (tx, ty, trow))
o_full = Array.uninit((n, n, n), np.float32, ctx=ctx)
o_half = o_full.slice(s_[: n // 2, :, :])
rowb = row.broadcast((n // 2, n, n))
xts, ys = x.t().slice(s_[::-2, :, :]), y.slice(s_[::2, :, :])
cases = {
    "combined y[::2]*row + x.t()[::-2]": (lambda: Zip.from_(o_half).and_(ys).and_(rowb).and_(xts).for_each("a=b*c+d"), 3 * n ** 3 * 2 + 4 * n),
    "x.t()[::-2] + y[::2]": (lambda: Zip.from_(o_half).and_(xts).and_(ys).for_each("a=b+c"), 3 * n ** 3 * 2),
    "x.t()[::2] copy": (lambda: Zip.from_(o_half).and_(x.t().slice(s_[::2, :, :])).for_each("a=b"), 2 * n ** 3 * 2),
    "x.t() + y (plain tile)": (lambda: Zip.from_(o_full).and_(x.t()).and_(y).for_each("a=b+c"), 3 * n ** 3 * 4),
}
for name, (fn, nbytes) in cases.items():
    med, best = _time(ctx, fn, 20, 3)
    print(f"{name:36s} {nbytes / med / 1e6:7.0f} GB/s  ({med * 1e3:.1f} us)  {nbytes / med / 1e6 / 6550:.3f} of peak", flush=True)
