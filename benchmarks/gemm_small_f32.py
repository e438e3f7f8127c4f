"""Small f32 general_mat_mul: the tcgen05 3xTF32 kernel (pack + GEMM) vs the FFMA tile kernel vs cuBLAS."""
import os
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ndarray_b200 as nd
from ndarray_b200 import Array
from benchmarks.gemm_probe import timeit
ctx = nd.default_context()
torch.backends.cuda.matmul.allow_tf32 = False
for n in (128, 192, 256, 384, 512, 640, 768, 1024):
    ta, tb, tc = (torch.randn(n, n, device="cuda") for _ in range(3))
    A, B, Cm = (Array.borrow(t.data_ptr(), (n, n), np.float32, "C", ctx, owner=t) for t in (ta, tb, tc))
    res = []
    for edge in ("128", "100000"):
        os.environ["NDB_GEMM_F32_MIN_CUBE"] = edge
        med, _ = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), 20, 5)
        res.append(med * 1e3)
    os.environ.pop("NDB_GEMM_F32_MIN_CUBE")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        torch.matmul(ta, tb, out=tc)
    e0.record()
    for _ in range(20):
        torch.matmul(ta, tb, out=tc)
    e1.record()
    torch.cuda.synchronize()
    print(f"n={n:5d}: tcgen05 {res[0]:7.1f} us   ffma tile {res[1]:7.1f} us   cuBLAS {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us", flush=True)
