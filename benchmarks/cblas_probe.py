"""cblas_dgemm / cblas_sgemm 8192^3 with pageable HOST operands: one-shot staging vs the row-block pipeline."""
import ctypes as C
import os
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ndarray_b200 as nd  # noqa: E402
L = nd.lib()
n = 8192
rng = np.random.default_rng(4)
for dt, name, cty in ((np.float64, "cblas_dgemm", C.c_double), (np.float32, "cblas_sgemm", C.c_float)):
    fn = getattr(L, name)
    fn.restype = None
    fn.argtypes = [C.c_int] * 6 + [cty, C.c_void_p, C.c_int, C.c_void_p, C.c_int, cty, C.c_void_p, C.c_int]
    ha, hb = rng.standard_normal((n, n), dtype=dt), rng.standard_normal((n, n), dtype=dt)
    hc = np.zeros((n, n), dtype=dt)
    for label, env in (("one shot", {"NDB_CBLAS_NO_PIPELINE": "1"}), ("pipelined", {}), ("pipelined, 8 blocks", {"NDB_CBLAS_BLOCKS": "8"}),
                       ("pipelined, 2 blocks", {"NDB_CBLAS_BLOCKS": "2"}), ("pipelined, 12 copy threads", {"NDB_COPY_THREADS": "12"}),
                       ("pipelined, 6 copy threads", {"NDB_COPY_THREADS": "6"})):
        os.environ.update(env)
        times = []
        for _ in range(4):
            t0 = time.perf_counter()
            fn(101, 111, 111, n, n, n, 1.0, ha.ctypes.data, n, hb.ctypes.data, n, 0.0, hc.ctypes.data, n)
            times.append(time.perf_counter() - t0)
        for k in env:
            os.environ.pop(k)
        best = min(times[1:])
        print(f"{name} {label:28s} {best * 1e3:7.1f} ms  {2.0 * n ** 3 / best / 1e12:6.2f} TFLOP/s", flush=True)
