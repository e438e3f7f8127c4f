"""SM clock and board power while a GEMM runs back to back (ours vs cuBLAS): is the tensor pipe power-capped?"""
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pynvml  # noqa: E402
import torch  # noqa: E402
import ndarray_b200 as nd  # noqa: E402
from ndarray_b200 import Array  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.stop = False
        self.clk, self.pw = [], []

    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            time.sleep(0.01)


def sustained(label, fn, sync, flops, seconds=3.0):
    fn()
    sync()
    s = Sampler()
    s.start()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(5):
            fn()
        sync()
        n += 5
    dt = time.perf_counter() - t0
    s.stop = True
    s.join()
    half = len(s.clk) // 2
    print(f"{label:28s} {flops * n / dt / 1e12:7.1f} TF  sm_mhz median {statistics.median(s.clk[half:]):.0f} "
          f"min {min(s.clk[half:])}  power median {statistics.median(s.pw[half:]):.0f} W max {max(s.pw):.0f} W", flush=True)


def main():
    ctx = nd.default_context()
    dev = torch.device("cuda", 0)
    n = 8192
    for name in sys.argv[1:] or ["f64", "f32"]:
        tdt, dt = (torch.float32, np.float32) if name == "f32" else (torch.float64, np.float64)
        ta, tb, tc = (torch.randn(n, n, device=dev, dtype=tdt) for _ in range(3))
        A, B, Cm = (Array.borrow(t.data_ptr(), (n, n), dt, "C", ctx, owner=t) for t in (ta, tb, tc))
        sustained(f"ours {name}", lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), ctx.sync, 2.0 * n ** 3)
        torch.backends.cuda.matmul.allow_tf32 = False
        sustained(f"cuBLAS {name}", lambda: torch.matmul(ta, tb, out=tc), torch.cuda.synchronize, 2.0 * n ** 3)
        if name == "f32":
            torch.backends.cuda.matmul.allow_tf32 = True
            sustained("cuBLAS 1xTF32", lambda: torch.matmul(ta, tb, out=tc), torch.cuda.synchronize, 2.0 * n ** 3)
            torch.backends.cuda.matmul.allow_tf32 = False


if __name__ == "__main__":
    main()
