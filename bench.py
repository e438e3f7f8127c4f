#!/usr/bin/env python
"""bench.py — headline measurement of the ndarray hot path on B200.

Workload (`config.workload`): BASELINE.json configs[1] — f32 512x512x512 non-contiguous views through Zip:
  (i)   out = x.t() + y                                   transposed operand   (1 610 612 736 algorithmic bytes)
  (ii)  half = x.slice(s![..;-2,..,..]) + y.slice(s![..;2,..,..])  negative step (805 306 368 B)
  (iii) out = x * row.broadcast((512,512,512))            stride-0 operand     (1 073 743 872 B)
One step = the three Zips once.  metric = algorithmic bytes moved per second (GB/s): every distinct
element read or written counts once.  Arrays are 512 MiB each, far beyond the 126 MB L2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]

N > 1 (launched under torchrun): each rank runs the same step on its own shard (weak scaling; Zip needs no
exchange), times are max-reduced over ranks, rank 0 prints the line.  `--impl reference` times the CPU
restatement of ndarray (oracle/, all host threads) on the same config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N3 = 512
BYTES_I = 3 * N3 ** 3 * 4
BYTES_II = 3 * (N3 // 2) * N3 * N3 * 4
BYTES_III = 2 * N3 ** 3 * 4 + N3 * 4
STEP_BYTES = BYTES_I + BYTES_II + BYTES_III
METRIC = "zip_views_hbm_gbps"
UNIT = "GB/s"
WORKLOAD = "C2: f32 512^3 Zip over views — x.t()+y, s![..;-2]+s![..;2], x*row.broadcast"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (a polling thread, ~1 kHz;
    nvidia-smi's 100 ms loop would miss a region that lasts tens of milliseconds)."""

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.err = None

    def _run(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:  # keep timing alive whatever NVML does
                self.err = repr(e)
                break
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


# -------------------------------------------------------------------------------------------------
def host_inputs(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, n, n), dtype=np.float32) * 2 - 1
    y = rng.random((n, n, n), dtype=np.float32) * 2 - 1
    row = rng.random(n, dtype=np.float32) * 2 - 1
    return x, y, row


def step_bytes(n):
    return 3 * n ** 3 * 4 + 3 * (n // 2) * n * n * 4 + 2 * n ** 3 * 4 + n * 4


def run_reference(args, rank, world):
    """The reference arm: ndarray's algorithm on the host cores (the oracle port; ndarray itself is Rust and
    the image has no rustc/cargo — DESIGN.md)."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    OP_ADD, OP_MUL = 2, 4
    # size the sample so one step takes a few seconds at most
    n = N3
    xs, ys, rs = host_inputs(128, 1)
    o = np.empty_like(xs)
    t0 = time.perf_counter()
    orc.zip_(OP_ADD, o, [xs.transpose(2, 1, 0), ys], threads=threads)
    t_small = time.perf_counter() - t0
    est = t_small * (N3 / 128) ** 3 * 2.2  # three cases ~ 2.2x the transposed one
    total_steps = args.steps + args.warmup
    while n > 128 and est * (n / N3) ** 3 * total_steps > 150.0:
        n //= 2
    x, y, row = host_inputs(n, 2)
    out = np.empty((n, n, n), np.float32)
    half = np.empty((n // 2, n, n), np.float32)
    rowb = np.broadcast_to(row, (n, n, n))

    def step():
        assert orc.zip_(OP_ADD, out, [x.transpose(2, 1, 0), y], threads=threads) == 0
        assert orc.zip_(OP_ADD, half, [x[::-1][::2], y[::2]], threads=threads) == 0
        assert orc.zip_(OP_MUL, out, [x, rowb], threads=threads) == 0

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = step_bytes(n) * args.steps / dt / 1e9
    sample = f"f32 {n}^3 (same three Zips), {args.steps} steps, oracle port of ndarray's Zip traversal, OpenMP static split along Zip::split's axis"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_n": n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(budget_s=20.0):
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    n = N3  # the bench workload itself (512^3), fewer passes: a smaller cube would sit in the CPU's last-level cache
    x, y, row = host_inputs(n, 3)
    out = np.empty((n, n, n), np.float32)
    half = np.empty((n // 2, n, n), np.float32)
    rowb = np.broadcast_to(row, (n, n, n))

    def step():
        orc.zip_(2, out, [x.transpose(2, 1, 0), y], threads=threads)
        orc.zip_(2, half, [x[::-1][::2], y[::2]], threads=threads)
        orc.zip_(4, out, [x, rowb], threads=threads)

    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < budget_s / 2 and reps < 60):
        step()
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": step_bytes(n) * reps / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"f32 {n}^3 views (same three Zips), {reps} passes in {dt:.1f} s, oracle/ndarray_oracle.cpp with OpenMP"}


def bind_to_gpu_numa_node(local: int) -> None:
    """Pin this process to the CPUs local to its GPU (PCIe root) so pinned host buffers are first-touched on that
    NUMA node; only matters for the end-to-end (H2D/D2H) leg when several ranks share the host."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = Path("/sys/bus/pci/devices") / bus.lower()[-12:] / "local_cpulist"
        cpus = set()
        for part in path.read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
    except Exception:
        pass


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import ndarray_b200 as nd
    from ndarray_b200 import Array, Zip, s_, Op, Axis
    from ndarray_b200._ffi import View, lib, check

    bind_to_gpu_numa_node(local)  # pinned staging buffers should live on the socket the GPU hangs off
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.setdefault("NDB_DEVICE", str(local))
    ctx = nd.Context(local)
    L = lib()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local)
        ctx.sync()

    # ---- inputs resident in HBM ---------------------------------------------------------------
    n = N3
    xh, yh, rowh = host_inputs(n, 100 + rank)
    x, y, row = Array.from_numpy(xh, ctx), Array.from_numpy(yh, ctx), Array.from_numpy(rowh, ctx)
    out = Array.uninit((n, n, n), np.float32, ctx=ctx)
    half = out.slice(s_[: n // 2, :, :])
    xt, xs, ys_, rowb = x.t(), x.slice(s_[::-2, :, :]), y.slice(s_[::2, :, :]), row.broadcast((n, n, n))
    z1 = Zip.from_(out).and_(xt).and_(y)
    z2 = Zip.from_(half).and_(xs).and_(ys_)
    z3 = Zip.from_(out).and_(x).and_(rowb)
    kernels = [nd.KERNEL_NAMES[z.plan(op).kernel] for z, op in ((z1, Op.ADD), (z2, Op.ADD), (z3, Op.MUL))]

    def step():
        z1.for_each("a=b+c")
        z2.for_each("a=b+c")
        z3.for_each("a=b*c")

    for _ in range(args.warmup):
        step()
    barrier()
    # timed region: events on the stream the kernels are launched on; per-kernel events ride along
    evs = [[ctx.event() for _ in range(4)] for _ in range(args.steps)]
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = L.ndb_launch_count()
    barrier()
    for k in range(args.steps):
        ctx.record(evs[k][0])
        z1.for_each("a=b+c")
        ctx.record(evs[k][1])
        z2.for_each("a=b+c")
        ctx.record(evs[k][2])
        z3.for_each("a=b*c")
        ctx.record(evs[k][3])
    barrier()
    launches = L.ndb_launch_count() - launches0
    clk = clocks.stop()
    total_ms = ctx.elapsed_ms(evs[0][0], evs[-1][3])
    per_kernel_ms = [statistics.mean(ctx.elapsed_ms(evs[k][j], evs[k][j + 1]) for k in range(args.steps)) for j in range(3)]
    if dist is not None:
        t = torch.tensor([total_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = STEP_BYTES * args.steps * world / (total_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peak, peak_src = peaks()
    case_bytes = [BYTES_I, BYTES_II, BYTES_III]
    dom = int(np.argmax(per_kernel_ms))
    achieved = case_bytes[dom] / (per_kernel_ms[dom] * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(["zip_tile", "zip_rows", "zip_rows"][dom])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": ["zip_tile<f32,ADD> (x.t()+y)", "zip_rows<f32,ADD,VEC> (s![..;-2])",
                                           "zip_rows<f32,MUL,VEC> (broadcast row)"][dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src,
                "per_case": {nm: {"ms": per_kernel_ms[j], "GB/s": case_bytes[j] / (per_kernel_ms[j] * 1e-3) / 1e9,
                                  "kernel_class": kernels[j]}
                             for j, nm in enumerate(["transposed", "neg_step", "broadcast_row"])}}

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------
    def pinned(arr):
        p = C.c_void_p()
        check(L.ndb_host_alloc(arr.nbytes, C.byref(p)))
        buf = np.ctypeslib.as_array((C.c_byte * arr.nbytes).from_address(p.value)).view(arr.dtype).reshape(arr.shape)
        buf[...] = arr
        return buf

    hx, hy, hrow = pinned(xh), pinned(yh), pinned(rowh)
    hout = pinned(np.zeros((n, n, n), np.float32))

    def hview(a, shape=None, strides=None, offset=0):
        v = View()
        shape = a.shape if shape is None else shape
        strides = [s // 4 for s in a.strides] if strides is None else strides
        v.ptr = a.ctypes.data + offset * 4
        v.ndim = len(shape)
        v.dtype = 0
        for i, (d, s) in enumerate(zip(shape, strides)):
            v.dim[i] = d
            v.stride[i] = s
        return v

    nn = n * n
    e2e_parts = [
        (Op.ADD, (View * 3)(hview(hout), hview(hx, (n, n, n), (1, n, nn)), hview(hy))),
        (Op.ADD, (View * 3)(hview(hout, (n // 2, n, n)), hview(hx, (n // 2, n, n), (-2 * nn, n, 1), (n - 1) * nn),
                            hview(hy, (n // 2, n, n), (2 * nn, n, 1)))),
        (Op.MUL, (View * 3)(hview(hout), hview(hx), hview(hrow, (n, n, n), (0, 0, 1)))),
    ]
    h2d = (2 * n ** 3 * 4) + (2 * (n // 2) * nn * 4) + (n ** 3 * 4 + n * 4)
    d2h = n ** 3 * 4 + (n // 2) * nn * 4 + n ** 3 * 4

    def e2e_step():
        for op, parts in e2e_parts:
            check(L.ndb_zip_host(ctx.handle, op, None, 3, parts))

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    es, ee = ctx.event(), ctx.event()
    ctx.record(es)
    for _ in range(e2e_steps):
        e2e_step()
    ctx.record(ee)
    barrier()
    e2e_ms = ctx.elapsed_ms(es, ee)
    if dist is not None:
        t = torch.tensor([e2e_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_val = STEP_BYTES * e2e_steps * world / (e2e_ms * 1e-3) / 1e9
    # the e2e result must equal the device result for case (iii), which ran last on both paths
    e2e_ok = bool(np.array_equal(hout[5], out.index_axis(0, 5).to_numpy()))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "shape": [n, n, n], "bytes_per_step": STEP_BYTES, "per_gpu_shard": "one 512^3 problem per rank",
                   "l2": "inputs 512 MiB each >> 126 MB L2, no flush needed", "kernel_classes": kernels},
        "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "api": "ndb_zip_host (pinned host buffers, chunked H2D/kernel/D2H overlap)",
                "matches_device_path": e2e_ok},
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_sample()
    if not args.no_extras:
        try:
            from benchmarks.extras import run_extras
            line["extra"] = run_extras(nd, ctx, torch, dist, rank, world, local, peak)
        except Exception as e:  # extras never break the headline line
            line["extra"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
