"""accumulate_axis_inplace (cumprod / running sum) throughput: f64 8192^2 and f32 512^3, every axis."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ndarray_b200 as nd
from ndarray_b200 import Array, Op
from benchmarks.gemm_probe import timeit

ctx = nd.default_context()
for shape, tdt, dt in (((8192, 8192), torch.float64, np.float64), ((512, 512, 512), torch.float32, np.float32)):
    t = torch.rand(*shape, device="cuda", dtype=tdt) * 0.01 + 0.995
    a = Array.borrow(t.data_ptr(), shape, dt, "C", ctx, owner=t)
    nbytes = 2 * t.numel() * t.element_size()  # one read + one write per element
    for axis in range(len(shape)):
        med, best = timeit(ctx, lambda: a.accumulate_axis_inplace(axis, Op.ADD), 5, 2)
        print(f"{np.dtype(dt).name} {shape} running sum along axis {axis}: {nbytes / med / 1e6:7.1f} GB/s ({med * 1e3:.1f} us)", flush=True)
