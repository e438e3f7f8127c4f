"""Secondary measurements printed under "extra" in bench.py's JSON line: the other BASELINE.json configs.

C1  f32 4096^2 contiguous  a = b*c + a  and sum_axis (buffers rotated so the 64 MiB arrays do not sit in L2)
C3  16384^2 f64 sum/mean/var_axis and i64 sum_axis / fold_axis(min), both axes
C4  general_mat_mul f64 / f32 at 8192^3 (C- and F-order operands)
C3  (N > 1) strong-scaled sharded sum_axis/var_axis(Axis(0)): NCCL vs peer windows vs the fused one-launch kernel
C4  (N > 1) row-sharded general_mat_mul with B broadcast over NCCL, one shot and pipelined in 8 row blocks
All timed with CUDA events on the context stream after warm-up; algorithmic bytes per SURVEY.md §8d.
"""
from __future__ import annotations

import statistics

import numpy as np


def _time(ctx, fn, iters, warm=2):
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


def run_extras(nd, ctx, torch, dist, rank, world, local, peak_gbs):
    from ndarray_b200 import Array, Axis, Zip
    dev = torch.device("cuda", local)
    out = {}

    def tarr(t, shape=None, order="C"):
        dt = {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32}[t.dtype]
        return Array.borrow(t.data_ptr(), tuple(t.shape) if shape is None else shape, dt, order, ctx, owner=t)

    def gbs(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9

    def rec(name, nbytes, med, best):
        out[name] = {"GB/s": round(gbs(nbytes, med), 1), "best_GB/s": round(gbs(nbytes, best), 1), "ms": round(med, 4),
                     "frac_of_measured_peak": round(gbs(nbytes, med) / peak_gbs, 3)}

    # ---- C1 ---------------------------------------------------------------------------------------
    n = 4096
    sets = []
    for _ in range(4):  # 4 sets x 3 arrays x 64 MiB = 768 MiB  >> L2
        ts = [torch.rand(n, n, device=dev) * 2 - 1 for _ in range(3)]
        sets.append([tarr(t) for t in ts])
    torch.cuda.synchronize(dev)
    it = [0]

    def c1_zip():
        a, b, c = sets[it[0] % 4]
        it[0] += 1
        Zip.from_(a).and_(b).and_(c).for_each("a=b*c+a")

    med, best = _time(ctx, c1_zip, 40, 4)
    rec("C1_zip_a=b*c+a_f32_4096^2", 4 * 4 * n * n, med, best)

    # one event pair around 12 launches (three passes over the four buffer sets): programmatic dependent launch overlaps each
    # kernel's launch with its predecessor's drain, which an event between every two launches prevents
    def c1_zip_x12():
        for _ in range(12):
            c1_zip()
    med, best = _time(ctx, c1_zip_x12, 10, 2)
    rec("C1_zip_a=b*c+a_f32_4096^2_12_back_to_back", 12 * 4 * 4 * n * n, med, best)
    import ctypes as C
    from ndarray_b200._ffi import lib, check
    o1 = Array.zeros((n,), np.float32, ctx=ctx)
    for ax in (0, 1):
        def c1_sum(ax=ax):
            a = sets[it[0] % 4][0]
            it[0] += 1
            a.sum_axis(Axis(ax))

        def c1_sum_into(ax=ax):
            a = sets[it[0] % 4][0]
            it[0] += 1
            check(lib().ndb_reduce_axis(ctx.handle, 0, ax, C.byref(a.v), C.byref(o1.v), 0.0))
        med, best = _time(ctx, c1_sum, 40, 4)
        rec(f"C1_sum_axis{ax}_f32_4096^2", 4 * n * n + 4 * n, med, best)
        med, best = _time(ctx, c1_sum_into, 40, 4)
        rec(f"C1_sum_axis{ax}_f32_4096^2_into_preallocated", 4 * n * n + 4 * n, med, best)

        # CUDA events tick in ~2 us steps on this box, too coarse for a 15 us kernel: also time 12 launches back to back
        # (three passes over the four 64 MiB buffer sets, so every input is cold in L2) between one event pair
        def c1_sum_x12(ax=ax):
            for _ in range(12):
                c1_sum_into(ax)
        med, best = _time(ctx, c1_sum_x12, 10, 2)
        rec(f"C1_sum_axis{ax}_f32_4096^2_into_preallocated_12_back_to_back", 12 * (4 * n * n + 4 * n), med, best)
    # the C1 loop as ONE CUDA graph (~10-40 us kernels): a = b*c + a, sum_axis(0), sum_axis(1) on each of the four
    # buffer sets = 12 kernels per replay, edges kept as programmatic dependent launches
    try:
        o0, o1v = Array.zeros((n,), np.float32, ctx=ctx), Array.zeros((n,), np.float32, ctx=ctx)

        def c1_body():
            for a, b, c in sets:
                Zip.from_(a).and_(b).and_(c).for_each("a=b*c+a")
                check(lib().ndb_reduce_axis(ctx.handle, 0, 0, C.byref(a.v), C.byref(o0.v), 0.0))
                check(lib().ndb_reduce_axis(ctx.handle, 0, 1, C.byref(a.v), C.byref(o1v.v), 0.0))
        step_bytes = 4 * (4 * 4 * n * n + 2 * (4 * n * n + 4 * n))
        med, best = _time(ctx, c1_body, 10, 2)
        rec("C1_step(zip+sum_axis0+sum_axis1)x4_stream_launches", step_bytes, med, best)
        with ctx.capture() as g:
            c1_body()
        med, best = _time(ctx, g.launch, 10, 2)
        rec("C1_step(zip+sum_axis0+sum_axis1)x4_one_cuda_graph", step_bytes, med, best)
        g.close()
    except Exception as e:  # noqa: BLE001
        out["C1_graph"] = {"error": repr(e)}
    del sets
    # ---- C2 variants: inner-axis step, map_collect (fresh output), all three views combined ------------
    n = 512
    from ndarray_b200 import s_, Op
    tx, ty = torch.rand(n, n, n, device=dev) * 2 - 1, torch.rand(n, n, n, device=dev) * 2 - 1
    trow = torch.rand(n, device=dev) * 2 - 1
    x, y, row = tarr(tx), tarr(ty), tarr(trow)
    o_full = Array.uninit((n, n, n), np.float32, ctx=ctx)
    o_in = o_full.slice(s_[:, :, : n // 2])
    med, best = _time(ctx, lambda: Zip.from_(o_in).and_(x.slice(s_[:, :, ::-2])).and_(y.slice(s_[:, :, ::2])).for_each("a=b+c"), 20, 3)
    rec("C2_inner_axis_step_s![..,..,..;-2]+s![..,..,..;2]", 3 * n ** 3 * 2, med, best)
    out["C2_inner_axis_step_s![..,..,..;-2]+s![..,..,..;2]"]["note"] = "stride-2 reads fetch 2x the algorithmic bytes in sectors (cap ~0.5)"
    med, best = _time(ctx, lambda: Zip.from_(x.t()).and_(y).map_collect(Op.ADD), 20, 3)
    rec("C2_transposed_via_map_collect(fresh output)", 3 * n ** 3 * 4, med, best)
    o_half = o_full.slice(s_[: n // 2, :, :])
    rowb = row.broadcast((n // 2, n, n))
    xts, ys = x.t().slice(s_[::-2, :, :]), y.slice(s_[::2, :, :])
    med, best = _time(ctx, lambda: Zip.from_(o_half).and_(ys).and_(rowb).and_(xts).for_each("a=b*c+d"), 20, 3)
    rec("C2_combined_y[::2]*row+x.t()[::-2]", 3 * n ** 3 * 2 + 4 * n, med, best)
    # ---- config 2's three view cases as ONE step (round 1's headline): device-resident, end to end through ndb_zip_host with
    #      pinned host buffers, and the CPU stand-in on the same views - the case where the host is slow (strided traversal) -----
    try:
        half = o_full.slice(s_[: n // 2, :, :])
        z1 = Zip.from_(o_full).and_(x.t()).and_(y)
        z2 = Zip.from_(half).and_(x.slice(s_[::-2, :, :])).and_(y.slice(s_[::2, :, :]))
        z3 = Zip.from_(o_full).and_(x).and_(row.broadcast((n, n, n)))
        c2_bytes = 3 * n ** 3 * 4 + 3 * (n // 2) * n * n * 4 + 2 * n ** 3 * 4 + n * 4

        def c2_step():
            z1.for_each("a=b+c")
            z2.for_each("a=b+c")
            z3.for_each("a=b*c")
        med, best = _time(ctx, c2_step, 20, 3)
        rec("C2_step_three_view_zips(x.t()+y, s![..;-2]+s![..;2], x*row)", c2_bytes, med, best)
        if rank == 0 and world == 1:
            from ndarray_b200._ffi import View
            import time as _t

            def pinned(shape):
                p = C.c_void_p()
                nb = int(np.prod(shape)) * 4
                check(lib().ndb_host_alloc(nb, C.byref(p)))
                return np.ctypeslib.as_array((C.c_byte * nb).from_address(p.value)).view(np.float32).reshape(shape)

            hx, hy, hrow, hout = pinned((n, n, n)), pinned((n, n, n)), pinned((n,)), pinned((n, n, n))
            rng = np.random.default_rng(3)
            hx[...] = rng.random((n, n, n), dtype=np.float32)
            hy[...] = rng.random((n, n, n), dtype=np.float32)
            hrow[...] = rng.random(n, dtype=np.float32)

            def hv(a, shape=None, strides=None, offset=0):
                v = View()
                shape = a.shape if shape is None else shape
                strides = [st // 4 for st in a.strides] if strides is None else strides
                v.ptr = a.ctypes.data + offset * 4
                v.ndim, v.dtype = len(shape), 0
                for i, (d, st) in enumerate(zip(shape, strides)):
                    v.dim[i], v.stride[i] = d, st
                return v

            nn = n * n
            parts = [(2, (View * 3)(hv(hout), hv(hx, (n, n, n), (1, n, nn)), hv(hy))),
                     (2, (View * 3)(hv(hout, (n // 2, n, n)), hv(hx, (n // 2, n, n), (-2 * nn, n, 1), (n - 1) * nn), hv(hy, (n // 2, n, n), (2 * nn, n, 1)))),
                     (4, (View * 3)(hv(hout), hv(hx), hv(hrow, (n, n, n), (0, 0, 1))))]

            def e2e():
                for op, pv in parts:
                    check(lib().ndb_zip_host(ctx.handle, op, None, 3, pv))
            e2e()
            t0 = _t.perf_counter()
            for _ in range(3):
                e2e()
            dt = (_t.perf_counter() - t0) / 3
            from oracle import oracle as orc
            orc.build()
            th = orc.max_threads()
            rowb = np.broadcast_to(hrow, (n, n, n))
            hhalf = hout[: n // 2]

            def cpu():
                orc.zip_(2, hout, [hx.transpose(2, 1, 0), hy], threads=th)
                orc.zip_(2, hhalf, [hx[::-1][::2], hy[::2]], threads=th)
                orc.zip_(4, hout, [hx, rowb], threads=th)
            cpu()
            t0 = _t.perf_counter()
            for _ in range(3):
                cpu()
            dc = (_t.perf_counter() - t0) / 3
            out["C2_step_end_to_end_vs_cpu"] = {
                "e2e_GB/s": round(c2_bytes / dt / 1e9, 1), "e2e_ms": round(dt * 1e3, 1), "h2d_bytes": 2 * n ** 3 * 4 + 2 * (n // 2) * nn * 4 + n ** 3 * 4 + n * 4,
                "d2h_bytes": n ** 3 * 4 + (n // 2) * nn * 4 + n ** 3 * 4, "cpu_stand_in_GB/s": round(c2_bytes / dc / 1e9, 1), "cpu_threads": th,
                "e2e_over_cpu": round(dc / dt, 2),
                "note": "strided views: the CPU traversal is cache-hostile, so even over PCIe the device path wins end to end; for the contiguous "
                        "config 5 the CPU streams at DDR5 speed and the link loses (bench e2e)"}
            for a_ in (hx, hy, hrow, hout):
                lib().ndb_host_free(C.c_void_p(a_.ctypes.data))
    except Exception as e:  # noqa: BLE001
        out["C2_step"] = {"error": repr(e)}
    del x, y, row, o_full, tx, ty
    # ---- C3 ---------------------------------------------------------------------------------------
    n = 16384
    tf = torch.randn(n, n, device=dev, dtype=torch.float64)
    f = tarr(tf)
    nb = 8 * n * n + 8 * n
    for ax in (0, 1):
        for name, fn in (("sum_axis", lambda ax=ax: f.sum_axis(Axis(ax))), ("mean_axis", lambda ax=ax: f.mean_axis(Axis(ax))),
                         ("var_axis", lambda ax=ax: f.var_axis(Axis(ax), 1.0))):
            med, best = _time(ctx, fn, 10, 2)
            rec(f"C3_{name}{ax}_f64_16384^2", nb, med, best)
    med, best = _time(ctx, lambda: f.sum(), 10, 2)
    rec("C3_sum_f64_16384^2", 8 * n * n, med, best)
    del f, tf
    ti = torch.randint(-2 ** 62, 2 ** 62, (n, n), device=dev, dtype=torch.int64)
    a = tarr(ti)
    for ax in (0, 1):
        med, best = _time(ctx, lambda ax=ax: a.sum_axis(Axis(ax)), 10, 2)
        rec(f"C3_sum_axis{ax}_i64_16384^2", nb, med, best)
        med, best = _time(ctx, lambda ax=ax: a.fold_axis(Axis(ax), np.iinfo(np.int64).max, "min"), 10, 2)
        rec(f"C3_fold_axis{ax}_min_i64_16384^2", nb, med, best)
    del a, ti
    torch.cuda.empty_cache()
    # ---- C4 ---------------------------------------------------------------------------------------
    n = 8192
    flops = 2.0 * n ** 3
    for tdt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        ta, tb, tc = (torch.randn(n, n, device=dev, dtype=tdt) for _ in range(3))
        for orders in ("CCC", "FFC"):
            A, B, Cm = (tarr(t, order=o) for t, o in zip((ta, tb, tc), orders))
            med, best = _time(ctx, lambda: nd.general_mat_mul(1.0, A, B, 0.0, Cm), 3, 1)
            out[f"C4_gemm_{name}_8192^3_{orders}"] = {"TFLOP/s": round(flops / (med * 1e-3) / 1e12, 2),
                                                      "best_TFLOP/s": round(flops / (best * 1e-3) / 1e12, 2), "ms": round(med, 3)}
        # library comparator (cuBLAS through torch), never on the product path
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.matmul(ta, tb)
        e0.record()
        for _ in range(3):
            torch.matmul(ta, tb)
        e1.record()
        torch.cuda.synchronize(dev)
        out[f"C4_comparator_cublas_{name}_8192^3"] = {"TFLOP/s": round(flops / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e12, 2)}
        if tdt == torch.float32:
            # what the tensor cores sustain for ONE tf32 MMA per logical MMA (cuBLAS TF32 mode): our 3xTF32 kernel
            # issues three, so its tensor-pipe rate is 3x its reported TFLOP/s
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.matmul(ta, tb)
            e0.record()
            for _ in range(3):
                torch.matmul(ta, tb)
            e1.record()
            torch.cuda.synchronize(dev)
            torch.backends.cuda.matmul.allow_tf32 = False
            out["C4_comparator_cublas_1xtf32_8192^3"] = {"TFLOP/s": round(flops / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e12, 2)}
        del ta, tb, tc
    torch.cuda.empty_cache()
    # ---- sharded axis reductions (C3 slabs) and the exchange alone; config 5 itself is bench.py's headline ------------
    if dist is not None and world > 1:
        import ctypes as C
        from ndarray_b200 import parallel as par
        from ndarray_b200.array import Red
        stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
        try:
            peers = par.PeerGroup(ctx, dist, slot_bytes=1 << 20)  # raises on ALL ranks or on none
        except Exception as e:  # noqa: BLE001
            peers = None
            out["peer_windows"] = {"error": repr(e)}

        def peer_section():
            # exchange latency alone for the C3 sum_axis(Axis(0)) vector (16384 f64 = 128 KiB)
            vec_t = torch.zeros(16384, device=dev, dtype=torch.float64)
            vec = tarr(vec_t)

            def nccl():
                with torch.cuda.stream(stream):
                    dist.all_reduce(vec_t)
            med_n, _ = _time(ctx, nccl, 50, 5)
            med_p, _ = _time(ctx, lambda: peers.allreduce_(vec), 50, 5)
            tn = torch.tensor([med_n, med_p], device=dev, dtype=torch.float64)
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
            out["allreduce_latency_128KiB"] = {"nccl_us": round(float(tn[0].item()) * 1e3, 1), "peer_memory_us": round(float(tn[1].item()) * 1e3, 1)}
            # C3 sharded, STRONG-scaled: sum_axis(Axis(0)) of this rank's [16384/N, 16384] f64 slab + exchange of the 128 KiB vector
            n = 16384
            rows = n // world
            slab_t = torch.randn(rows, n, device=dev, dtype=torch.float64)
            slab = tarr(slab_t)

            def two_launch():
                part = slab.sum_axis(0)
                return peers.allreduce_(part)

            fused_ok = par._fused_axis0(slab, Red.SUM, 0.0, n, peers) is not None  # same answer on every rank
            cases = [("nccl", lambda: par.sum_axis0(slab, dist)), ("peer", two_launch), ("var_nccl", lambda: par.var_axis0(slab, 1.0, dist))]
            if fused_ok:
                cases += [("peer_fused_one_launch", lambda: par._fused_axis0(slab, Red.SUM, 0.0, n, peers)),
                          ("var_peer_fused_one_launch", lambda: par._fused_axis0(slab, Red.VAR, 1.0, n, peers))]
            for label, fn in cases:
                med_s, _ = _time(ctx, fn, 20, 3)
                ts = torch.tensor([med_s], device=dev, dtype=torch.float64)
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
                out[f"C3_sum_axis0_sharded_f64_{label}"] = {"GB/s_aggregate": round(gbs(8 * n * n + 8 * n, float(ts.item())), 1),
                                                           "ms": round(float(ts.item()), 4), "scaling": "strong: 16384^2 f64 over all ranks"}
            # what actually crosses NVLink for the fused kernel: NVML's per-GPU NVLink data counters (KiB, all links) around
            # 200 launches - the states are tiny next to the 2 GiB / N each rank streams from its own HBM
            try:
                import pynvml as nv
                nv.nvmlInit()
                h = nv.nvmlDeviceGetHandleByIndex(local)
                ids = [getattr(nv, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", 138), getattr(nv, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", 139)]

                def nvlink_kib():
                    vals = nv.nvmlDeviceGetFieldValues(h, ids)
                    return [int(v.value.ullVal) for v in vals]

                if fused_ok:
                    ctx.sync()
                    dist.barrier()
                    t0 = nvlink_kib()
                    reps = 200
                    for _ in range(reps):
                        par._fused_axis0(slab, Red.VAR, 1.0, n, peers)
                    ctx.sync()
                    dist.barrier()
                    t1 = nvlink_kib()
                    tx, rx = (t1[0] - t0[0]) * 1024 / reps, (t1[1] - t0[1]) * 1024 / reps
                    expect = (world - 1) * n * 24  # Welford (n, mean, M2) per column to every other rank
                    rec_nv = {"expected_payload": expect, "hbm_bytes_read_per_rank": 8 * rows * n,
                              "source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX (all links of this GPU), rank 0, 200 launches"}
                    if tx == 0 and rx == 0:
                        rec_nv["counters"] = "unavailable: the NVML NVLink throughput fields read 0 on this (virtualised) box"
                    else:
                        rec_nv.update({"tx": round(tx), "rx": round(rx)})
                    out["C3_var_axis0_fused_nvlink_bytes_per_launch"] = rec_nv
            except Exception as e:  # noqa: BLE001
                out["C3_var_axis0_fused_nvlink_bytes_per_launch"] = {"error": repr(e)}
            ctx.sync()
            dist.barrier()
            peers.close()

        if peers is not None:
            peer_section()
        torch.cuda.empty_cache()
    # ---- C4 at N GPUs: A and C row-sharded, B broadcast from rank 0 EVERY step.  Weak (8192 rows per rank, as in round 1) and
    #      STRONG (the 8192^3 problem of config 4: 8192 / N rows per rank), the latter also through the C ABI alone
    #      (ndb_gemm_rowsharded: the library's own NCCL communicator, no torch collective) ---------------------------------
    if dist is not None and world > 1:
        from ndarray_b200 import parallel as par
        n = 8192
        comm = None
        try:
            comm = par.Comm(ctx, dist)
        except Exception as e:  # noqa: BLE001
            out["ndb_comm"] = {"error": repr(e)}
        for tdt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
            ta, tb, tc = (torch.randn(n, n, device=dev, dtype=tdt) for _ in range(3))
            A, B, Cm = (tarr(t) for t in (ta, tb, tc))
            mib = n * n * ta.element_size() >> 20
            for chunks in (1, 8):
                med, best = _time(ctx, lambda: par.general_mat_mul_rowsharded(1.0, A, B, 0.0, Cm, dist, 0, chunks), 3, 1)
                t = torch.tensor([med], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                out[f"C4_gemm_{name}_rowsharded_{world}gpu_bcast_chunks{chunks}"] = {
                    "TFLOP/s_aggregate": round(world * 2.0 * n ** 3 / (float(t.item()) * 1e-3) / 1e12, 1), "ms": round(float(t.item()), 3),
                    "scaling": "weak: 8192 rows per rank", "collective": "ncclBroadcast of B (%d MiB) inside the timed step" % mib}
            rows = n // world
            As, Cs = A.slice_axis(0, nd.Slice(0, rows, 1)), Cm.slice_axis(0, nd.Slice(0, rows, 1))
            variants = [("torch_nccl", lambda: par.general_mat_mul_rowsharded(1.0, As, B, 0.0, Cs, dist, 0, 1))]
            if comm is not None:
                variants.append(("c_abi_ndb_gemm_rowsharded", lambda: par.gemm_rowsharded(1.0, As, B, 0.0, Cs, 0)))
            for label, fn in variants:
                med, best = _time(ctx, fn, 5, 2)
                t = torch.tensor([med], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                out[f"C4_gemm_{name}_8192^3_strong_{world}gpu_{label}"] = {
                    "TFLOP/s_aggregate": round(2.0 * n ** 3 / (float(t.item()) * 1e-3) / 1e12, 1), "ms": round(float(t.item()), 3),
                    "scaling": "strong: 8192^3 in all, %d rows of A and C per rank" % rows,
                    "collective": "broadcast of B (%d MiB) inside the timed step" % mib}
            del ta, tb, tc
        if comm is not None:
            ctx.sync()
            dist.barrier()
            comm.close()
        torch.cuda.empty_cache()
    # ---- C4 through ndarray's own plugin boundary: cblas_{d,s}gemm with HOST pointers (what `features = ["blas"]` calls;
    #      pageable host memory like ndarray's Vec, H2D of A and B and D2H of C inside the timed call) -----------------
    if rank == 0 and world == 1:
        try:
            import ctypes as C
            import time as _t
            L = nd.lib()
            n = 8192
            rng = np.random.default_rng(4)
            for dt, fn_name, cty in ((np.float64, "cblas_dgemm", C.c_double), (np.float32, "cblas_sgemm", C.c_float)):
                fn = getattr(L, fn_name)
                fn.restype = None
                fn.argtypes = [C.c_int] * 6 + [cty, C.c_void_p, C.c_int, C.c_void_p, C.c_int, cty, C.c_void_p, C.c_int]
                ha = rng.standard_normal((n, n), dtype=dt)
                hb = rng.standard_normal((n, n), dtype=dt)
                hc = np.zeros((n, n), dtype=dt)
                times = []
                for _ in range(3):
                    t0 = _t.perf_counter()
                    fn(101, 111, 111, n, n, n, 1.0, ha.ctypes.data, n, hb.ctypes.data, n, 0.0, hc.ctypes.data, n)
                    times.append(_t.perf_counter() - t0)
                best = min(times[1:])
                out[f"C4_{fn_name}_8192^3_host_pointers_end_to_end"] = {
                    "TFLOP/s": round(2.0 * n ** 3 / best / 1e12, 2), "ms": round(best * 1e3, 1),
                    "host_bytes_moved": 3 * n * n * np.dtype(dt).itemsize, "note": "pageable host buffers, staged H2D/D2H inside the call"}
                del ha, hb, hc
        except Exception as e:
            out["C4_cblas_end_to_end"] = {"error": repr(e)}
    if rank == 0 and world == 1:
        try:
            out["cpu_stand_in"] = cpu_stand_ins()
        except Exception as e:  # never break the bench line
            out["cpu_stand_in"] = {"error": repr(e)}
    return out


def cpu_stand_ins():
    """The reference's CPU path for the other configs, timed on this box's host cores on bounded samples (the
    in-image stand-in of BASELINE.md §3: the oracle port of ndarray's algorithms; OpenBLAS through numpy for GEMM,
    which is what ndarray's `blas` feature would call).  ndarray has no parallel axis reductions, so those are serial."""
    import time
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    res = {"threads": threads}
    rng = np.random.default_rng(0)

    def best(fn, reps=3):
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            t.append(time.perf_counter() - t0)
        return min(t)

    n = 4096
    a, b, c = (rng.random((n, n), dtype=np.float32) for _ in range(3))
    dt = best(lambda: orc.zip_(7, a, [b, c, a], threads=threads))
    res["C1_zip_a=b*c+a_f32_4096^2"] = {"GB/s": round(16 * n * n / dt / 1e9, 1), "threads": threads, "kind": "oracle port, OpenMP"}
    for ax in (0, 1):
        dt = best(lambda: orc.reduce_axis(0, a, ax))
        res[f"C1_sum_axis{ax}_f32_4096^2"] = {"GB/s": round(4 * n * n / dt / 1e9, 1), "threads": 1, "kind": "oracle port (serial, as in ndarray)"}
    f = rng.standard_normal((n, n))
    for name, red in (("sum_axis", 0), ("var_axis", 5)):
        for ax in (0, 1):
            dt = best(lambda: orc.reduce_axis(red, f, ax, 1.0), 2)
            res[f"C3_{name}{ax}_f64_sample_4096^2"] = {"GB/s": round(8 * n * n / dt / 1e9, 1), "threads": 1, "kind": "oracle port (serial), 1/16 of config 3"}
    m = 2048
    for dtp, name in ((np.float64, "f64"), (np.float32, "f32")):
        A, B = rng.standard_normal((m, m)).astype(dtp), rng.standard_normal((m, m)).astype(dtp)
        A @ B
        dt = best(lambda: A @ B, 3)
        res[f"C4_gemm_{name}_sample_2048^3"] = {"TFLOP/s": round(2 * m ** 3 / dt / 1e12, 3), "kind": "OpenBLAS via numpy (what ndarray's blas feature calls)"}
    return res
