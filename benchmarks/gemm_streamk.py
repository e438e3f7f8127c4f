"""f64 general_mat_mul with and without stream-K (NDB_GEMM_F64_STREAMK) against cuBLAS, at sizes whose last wave is ragged."""
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
dev = torch.device("cuda", 0)
for (m, k, n) in ((2048, 2048, 2048), (2560, 2560, 2560), (1792, 4096, 1792), (2048, 8192, 2048), (2304, 2048, 2304), (3072, 3072, 3072), (1536, 2048, 1536),
                  (1024, 8192, 8192)):  # the last one is one rank's share of config 4 on 8 GPUs (512 tiles = 3.46 waves)
    ta = torch.randn(m, k, device=dev, dtype=torch.float64)
    tb = torch.randn(k, n, device=dev, dtype=torch.float64)
    tc = torch.zeros(m, n, device=dev, dtype=torch.float64)
    A = Array.borrow(ta.data_ptr(), (m, k), np.float64, "C", ctx, owner=ta)
    B = Array.borrow(tb.data_ptr(), (k, n), np.float64, "C", ctx, owner=tb)
    Cm = Array.borrow(tc.data_ptr(), (m, n), np.float64, "C", ctx, owner=tc)
    res = {}
    for mode in ("0", "1", None):
        if mode is None:
            os.environ.pop("NDB_GEMM_F64_STREAMK", None)
        else:
            os.environ["NDB_GEMM_F64_STREAMK"] = mode
        med, _ = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), 20, 5)
        res[mode] = med
    os.environ.pop("NDB_GEMM_F64_STREAMK", None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        torch.matmul(ta, tb, out=tc)
    e0.record()
    for _ in range(20):
        torch.matmul(ta, tb, out=tc)
    e1.record()
    torch.cuda.synchronize()
    cb = e0.elapsed_time(e1) / 20
    fl = 2.0 * m * k * n
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    print(f"f64 {m}x{k}x{n} ({tiles} tiles = {tiles / 148:.2f} waves): whole tiles {res['0'] * 1e3:7.1f} us ({fl / res['0'] / 1e9:5.2f} TF)  "
          f"stream-K {res['1'] * 1e3:7.1f} us ({fl / res['1'] / 1e9:5.2f} TF)  default {res[None] * 1e3:7.1f} us  "
          f"cuBLAS {cb * 1e3:7.1f} us ({fl / cb / 1e9:5.2f} TF)  default/cuBLAS {cb / res[None]:.3f}", flush=True)
