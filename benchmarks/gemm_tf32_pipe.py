"""Is the 3xTF32 kernel at the tensor roofline at every size, or only at 8192^3?  Its tensor-pipe rate (3 MMAs per logical
MMA: 3 x effective TFLOP/s, packing pre-pass included) next to cuBLAS in 1xTF32 mode (one MMA per logical MMA)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ndarray_b200 as nd
from ndarray_b200 import Array
from benchmarks.gemm_probe import timeit
ctx = nd.default_context()
for n in (1024, 2048, 3072, 4096, 6144, 8192):
    ta, tb, tc = (torch.randn(n, n, device="cuda") for _ in range(3))
    A, B, Cm = (Array.borrow(t.data_ptr(), (n, n), np.float32, "C", ctx, owner=t) for t in (ta, tb, tc))
    med, _ = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), 6, 2)
    torch.backends.cuda.matmul.allow_tf32 = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        torch.matmul(ta, tb, out=tc)
    e0.record()
    for _ in range(6):
        torch.matmul(ta, tb, out=tc)
    e1.record()
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    cb = e0.elapsed_time(e1) / 6
    fl = 2.0 * n ** 3
    print(f"n={n:5d}: ours {fl / med / 1e9:6.1f} TF effective = {3 * fl / med / 1e9:6.1f} on the tensor pipe   cuBLAS 1xTF32 {fl / cb / 1e9:6.1f} TF   pipe ratio {3 * cb / med:4.2f}", flush=True)
