"""general_mat_mul across sizes, ours vs cuBLAS (torch.matmul), f64 and f32."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ndarray_b200 as nd
from ndarray_b200 import Array
from benchmarks.gemm_probe import timeit

ctx = nd.default_context()
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
for name, tdt, dt in (("f64", torch.float64, np.float64), ("f32", torch.float32, np.float32)):
    for n in (128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096):
        ta, tb, tc = (torch.randn(n, n, device=dev, dtype=tdt) for _ in range(3))
        A, B, Cm = (Array.borrow(t.data_ptr(), (n, n), dt, "C", ctx, owner=t) for t in (ta, tb, tc))
        med, _ = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), 10, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            torch.matmul(ta, tb, out=tc)
        e0.record()
        for _ in range(10):
            torch.matmul(ta, tb, out=tc)
        e1.record()
        torch.cuda.synchronize()
        cb = e0.elapsed_time(e1) / 10
        fl = 2.0 * n ** 3
        print(f"{name} n={n:5d}: ours {fl / med / 1e9:7.2f} TF ({med * 1e3:8.1f} us)   cuBLAS {fl / cb / 1e9:7.2f} TF ({cb * 1e3:8.1f} us)   ratio {cb / med:5.2f}", flush=True)
