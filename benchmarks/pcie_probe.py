"""Host<->device copy ceilings on this box (pinned memory), to judge the e2e path against."""
import ctypes as C
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200._ffi import lib, check  # noqa: E402


def main():
    ctx = nd.default_context()
    L = lib()
    n = 1 << 30
    h = C.c_void_p()
    check(L.ndb_host_alloc(n, C.byref(h)))
    h2 = C.c_void_p()
    check(L.ndb_host_alloc(n, C.byref(h2)))
    d = C.c_void_p()
    check(L.ndb_alloc(ctx.handle, n, C.byref(d)))
    d2 = C.c_void_p()
    check(L.ndb_alloc(ctx.handle, n, C.byref(d2)))
    for name, fn in (("H2D 1 GiB", lambda: L.ndb_h2d(ctx.handle, d, h, n)), ("D2H 1 GiB", lambda: L.ndb_d2h(ctx.handle, h2, d2, n))):
        fn(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        ctx.sync()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name}: {n / dt / 1e9:.1f} GB/s")
    # both directions at once on two streams
    import torch
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    hp = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    hq = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    dp = torch.empty(n, dtype=torch.uint8, device="cuda")
    dq = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            dp.copy_(hp, non_blocking=True)
        with torch.cuda.stream(s2):
            hq.copy_(dq, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"H2D + D2H concurrently: {n / dt / 1e9:.1f} GB/s each direction")


if __name__ == "__main__":
    main()
