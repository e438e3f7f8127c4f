"""GEMM timing probe (random normal inputs, like the bench) under a few env settings."""
import os
import sys
import statistics
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array  # noqa: E402


def timeit(ctx, fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    evs = [(ctx.event(), ctx.event()) for _ in range(iters)]
    for a, b in evs:
        ctx.record(a)
        fn()
        ctx.record(b)
    ctx.sync()
    ms = [ctx.elapsed_ms(a, b) for a, b in evs]
    return statistics.median(ms), min(ms)


def main():
    ctx = nd.default_context()
    dev = torch.device("cuda", 0)
    which = sys.argv[1:] or ["f32", "f64"]
    for name in which:
        tdt, dt = (torch.float32, np.float32) if name == "f32" else (torch.float64, np.float64)
        for n in (4096, 8192):
            ta, tb, tc = (torch.randn(n, n, device=dev, dtype=tdt) for _ in range(3))
            A, B, Cm = (Array.borrow(t.data_ptr(), (n, n), dt, "C", ctx, owner=t) for t in (ta, tb, tc))
            Bf = Array.borrow(tb.data_ptr(), (n, n), dt, "F", ctx, owner=tb)
            for envs in ([{}] + ([{"NDB_GEMM_DEBUG": g} for g in ("1", "2")] if name == "f32" else [])):
                for k in ("NDB_GEMM_GROUP_M", "NDB_GEMM_CHUNK_KB", "NDB_GEMM_DEBUG"):
                    os.environ.pop(k, None)
                os.environ.update(envs)
                med, best = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm))
                medf, _ = timeit(ctx, lambda: nd.general_mat_mul(1.0, A, Bf, 0.0, Cm))
                print(f"{name} n={n} {envs}: CC {med:7.3f} ms {2*n**3/med/1e9:7.1f} TF | B=F {medf:7.3f} ms {2*n**3/medf/1e9:7.1f} TF", flush=True)
            # accuracy vs float64 cuBLAS on a sample of rows
            nd.general_mat_mul(1.0, A, B, 0.0, Cm)
            ctx.sync()
            ref = (ta[:64].double() @ tb.double())
            err = ((tc[:64].double() - ref).abs().max() / ref.abs().max()).item()
            print(f"   rel-norm err vs f64 reference on 64 rows: {err:.3e}")


if __name__ == "__main__":
    main()
