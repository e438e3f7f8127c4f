"""torchrun --nproc-per-node N benchmarks/dist_check.py — the multi-GPU protocol of ndarray_b200/parallel.py on
real GPUs: every rank holds the same host data, takes its Axis(0) slab, and the sharded results must equal the
unsharded numpy results (integers exactly, floats within the north-star tolerance).  Covers the NCCL exchange, the
peer-window exchange (csrc/peer.cu) and the fused one-launch reductions (red_cols<PEER>), including the shapes that
used to be able to split the ranks: fewer rows than ranks (empty shards), a single row per rank, a shard that is
not 16-byte aligned on some ranks only.

`run_checks` is also called by bench.py at N > 1 (before the timed region), so the SCALE record carries
`"dist_parity": "ok"`.  Checks never raise mid-way (a rank that bailed out would leave the others in a collective):
failures are collected and gathered at the end."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def run_checks(ctx, dist, rank, world, gemm: bool = True, peers=None):
    """Returns (n_checks, failures) with the failures of ALL ranks (gathered)."""
    from ndarray_b200 import Array, Zip, parallel as par
    from ndarray_b200.array import Red
    fails, count = [], [0]

    def ok(name, cond, detail=""):
        count[0] += 1
        if not bool(cond):
            fails.append(f"rank {rank}: {name} {detail}")

    def close(name, got, want, tol):
        err = float(np.max(np.abs(np.asarray(got, dtype=np.float64) - np.asarray(want, dtype=np.float64)))) if np.size(want) else 0.0
        ok(name, err <= tol, f"err {err:.3e} > tol {tol:.3e}")

    rng = np.random.default_rng(7)
    n0 = 1001
    a = rng.integers(-2 ** 62, 2 ** 62, size=(n0, 300), dtype=np.int64)
    f = rng.standard_normal((n0, 300))
    g = rng.uniform(-1, 1, (n0, 300)).astype(np.float32)
    lo, hi = par.shard_bounds(n0, rank, world)
    da, df, dg = (Array.from_numpy(x, ctx) for x in (a, f, g))
    sa, sf, sg = (par.shard_axis0(x, rank, world) for x in (da, df, dg))
    ok("shard rows", sa.shape[0] == hi - lo)
    # elementwise on the shard: no exchange
    out = Array.uninit(sg.shape, np.float32, ctx=ctx)
    Zip.from_(out).and_(sg).and_(sg).for_each("a=b*c")
    ok("sharded Zip bit-exact", np.array_equal(out.to_numpy(), g[lo:hi] * g[lo:hi]))
    tol = 1e-12 * np.sqrt(n0)
    own_peers = peers is None
    if own_peers:
        peers = par.PeerGroup(ctx, dist, slot_bytes=1 << 20)
    with np.errstate(over="ignore"):
        for label, pg, reps in (("nccl", None, 1), ("peer", peers, 3)):  # repeated: both slot sets and the counters keep working
            for rep in range(reps):
                ok(f"{label} global_sum i64", par.global_sum(sa, dist, pg) == a.sum(dtype=np.int64))
                ok(f"{label} sum_axis0 i64", np.array_equal(par.sum_axis0(sa, dist, pg, n0).to_numpy(), a.sum(axis=0, dtype=np.int64)))
                close(f"{label} global_sum f64", par.global_sum(sf, dist, pg), f.sum(), tol * 300 * 8)
                close(f"{label} sum_axis0 f64", par.sum_axis0(sf, dist, pg, n0).to_numpy(), f.sum(axis=0), tol * 8)
                close(f"{label} mean_axis0", par.mean_axis0(sf, n0, dist, pg).to_numpy(), f.mean(axis=0), tol)
                close(f"{label} var_axis0", par.var_axis0(sf, 1.0, dist, peers=pg, n_global=n0).to_numpy(), f.var(axis=0, ddof=1), tol * 8)
                close(f"{label} var_axis0 unfused", par.var_axis0(sf, 1.0, dist, peers=pg).to_numpy(), f.var(axis=0, ddof=1), tol * 8)
                close(f"{label} std_axis0", par.var_axis0(sf, 0.0, dist, True, pg, n0).to_numpy(), f.std(axis=0), tol * 8)
        # the fused one-launch path must actually be taken for a wide slab, and agree with numpy
        wide = rng.standard_normal((n0, 4096))
        wl = par.shard_axis0(Array.from_numpy(wide, ctx), rank, world)
        for name, red, ddof, want in (("sum", Red.SUM, 0.0, wide.sum(axis=0)), ("var", Red.VAR, 1.0, wide.var(axis=0, ddof=1)),
                                      ("mean", Red.MEAN, 0.0, wide.mean(axis=0)), ("std", Red.STD, 0.0, wide.std(axis=0))):
            got = par._fused_axis0(wl, red, ddof, n0, peers)
            ok(f"fused {name} taken", got is not None)
            if got is not None:
                close(f"fused {name}", got.to_numpy(), want, tol * 8)
        wi = rng.integers(-2 ** 62, 2 ** 62, size=(n0, 2048), dtype=np.int64)
        fi = par._fused_axis0(par.shard_axis0(Array.from_numpy(wi, ctx), rank, world), Red.SUM, 0.0, n0, peers)
        ok("fused i64 sum", fi is not None and np.array_equal(fi.to_numpy(), wi.sum(axis=0, dtype=np.int64)))
        for name, red, want in (("min", Red.MIN, wi.min(axis=0)), ("max", Red.MAX, wi.max(axis=0))):
            got = par._fused_axis0(par.shard_axis0(Array.from_numpy(wi, ctx), rank, world), red, 0.0, n0, peers)
            ok(f"fused i64 {name}", got is not None and np.array_equal(got.to_numpy(), want))
        # rank-local accidents must not split the ranks: fewer rows than ranks (empty shards on some), one row per rank,
        # and a slab that starts 8 bytes off a 16-byte boundary on the odd ranks only
        for rows in (max(1, world // 2), world, world + 1):
            tiny = rng.standard_normal((rows, 1024))
            tl = par.shard_axis0(Array.from_numpy(tiny, ctx), rank, world)
            got = par._fused_axis0(tl, Red.SUM, 0.0, rows, peers)
            ok(f"fused sum, {rows} rows on {world} ranks, taken", got is not None)
            if got is not None:
                close(f"fused sum, {rows} rows on {world} ranks", got.to_numpy(), tiny.sum(axis=0), tol * 8)
            got = par.sum_axis0(tl, dist, peers, rows)
            close(f"sum_axis0, {rows} rows on {world} ranks", got.to_numpy(), tiny.sum(axis=0), tol * 8)
            if rows > 1:
                got = par.var_axis0(tl, 1.0, dist, peers=peers, n_global=rows)
                close(f"var_axis0, {rows} rows on {world} ranks", got.to_numpy(), tiny.var(axis=0, ddof=1), tol * 8)
        shard = wide[lo:hi]
        if rank % 2 == 1:
            flat = Array.uninit((shard.size + 1,), np.float64, ctx=ctx)
            mis = flat.raw_view_with(1, shard.shape, (shard.shape[1], 1))  # same strides, base pointer off by one element
            mis.assign(Array.from_numpy(shard, ctx))
        else:
            mis = Array.from_numpy(shard, ctx)
        got = par._fused_axis0(mis, Red.VAR, 1.0, n0, peers)
        ok("fused var, odd ranks misaligned, taken", got is not None)
        if got is not None:
            close("fused var, odd ranks misaligned", got.to_numpy(), wide.var(axis=0, ddof=1), tol * 8)
        # peer exchange and NCCL agree to the bit for integers
        ok("peer == nccl (i64)", np.array_equal(par.sum_axis0(sa, dist, peers, n0).to_numpy(), par.sum_axis0(sa, dist).to_numpy()))
    gathered = peers.allgather(Array.from_numpy(np.full(5, rank, dtype=np.int32), ctx)).to_numpy()
    ok("peer allgather", np.array_equal(gathered, np.repeat(np.arange(world, dtype=np.int32)[:, None], 5, axis=1)))
    mx = peers.allreduce_(Array.from_numpy(np.array([rank, -rank], dtype=np.float64), ctx), Red.MAX).to_numpy()
    ok("peer allreduce max", np.array_equal(mx, [world - 1, 0.0]))
    if gemm:
        # general_mat_mul: A and C row-sharded, B broadcast from rank 0 (one shot and pipelined in 4 row blocks of B)
        for dt, tol_mm in ((np.float64, 1e-12), (np.float32, 1e-5)):
            A = rng.standard_normal((n0, 512)).astype(dt)
            B = rng.standard_normal((512, 300)).astype(dt)
            C0 = rng.standard_normal((n0, 300)).astype(dt)
            ref = (-2.3 * (A.astype(np.float64) @ B.astype(np.float64)) + np.pi * C0)[lo:hi]
            for chunks in (1, 4):
                dA = par.shard_axis0(Array.from_numpy(A, ctx), rank, world)
                dC = par.shard_axis0(Array.from_numpy(C0, ctx), rank, world)
                dB = Array.from_numpy(B if rank == 0 else np.zeros_like(B), ctx)  # only rank 0 holds B
                par.general_mat_mul_rowsharded(-2.3, dA, dB, np.pi, dC, dist, src=0, k_chunks=chunks)
                err = np.max(np.abs(dC.to_numpy() - ref)) / np.max(np.abs(ref)) if hi > lo else 0.0
                ok(f"row-sharded gemm {np.dtype(dt).name} chunks={chunks}", err <= tol_mm * np.sqrt(512), f"rel err {err:.3e}")
            if hasattr(par, "general_mat_mul_rowsharded_peer"):
                dA = par.shard_axis0(Array.from_numpy(A, ctx), rank, world)
                dC = par.shard_axis0(Array.from_numpy(C0, ctx), rank, world)
                dB = Array.from_numpy(B if rank == 0 else np.zeros_like(B), ctx)
                par.general_mat_mul_rowsharded_peer(-2.3, dA, dB, np.pi, dC, peers, src=0)
                err = np.max(np.abs(dC.to_numpy() - ref)) / np.max(np.abs(ref)) if hi > lo else 0.0
                ok(f"row-sharded gemm over peer windows {np.dtype(dt).name}", err <= tol_mm * np.sqrt(512), f"rel err {err:.3e}")
    # ---- the sharded entry points of the C ABI (csrc/dist.cu): library-owned NCCL communicator, no Python collectives ----
    import ndarray_b200 as nd
    for label, cctx in (("nccl only", nd.Context(ctx.device)), ("nccl + peer windows", ctx)):
        comm = par.Comm(cctx, dist)
        ca, cf = Array.from_numpy(a, cctx), Array.from_numpy(f, cctx)
        csa, csf = par.shard_axis0(ca, rank, world), par.shard_axis0(cf, rank, world)
        with np.errstate(over="ignore"):
            ok(f"C ABI [{label}] sum() i64", par.reduce_all_sharded(csa) == a.sum(dtype=np.int64))
            ok(f"C ABI [{label}] sum_axis0 i64", np.array_equal(par.reduce_axis0_sharded(csa, Red.SUM, n0).to_numpy(), a.sum(axis=0, dtype=np.int64)))
            ok(f"C ABI [{label}] fold_axis0 min i64", np.array_equal(par.reduce_axis0_sharded(csa, Red.MIN, n0).to_numpy(), a.min(axis=0)))
            ok(f"C ABI [{label}] min() / max() i64", par.reduce_all_sharded(csa, Red.MIN) == a.min() and par.reduce_all_sharded(csa, Red.MAX) == a.max())
        ok(f"C ABI [{label}] min() / max() f64", par.reduce_all_sharded(csf, Red.MIN) == f.min() and par.reduce_all_sharded(csf, Red.MAX) == f.max())
        close(f"C ABI [{label}] sum() f64", par.reduce_all_sharded(csf), f.sum(), tol * 300 * 8)
        close(f"C ABI [{label}] sum_axis0 f64", par.reduce_axis0_sharded(csf, Red.SUM, n0).to_numpy(), f.sum(axis=0), tol * 8)
        close(f"C ABI [{label}] mean_axis0", par.reduce_axis0_sharded(csf, Red.MEAN, n0).to_numpy(), f.mean(axis=0), tol)
        close(f"C ABI [{label}] var_axis0", par.reduce_axis0_sharded(csf, Red.VAR, n0, 1.0).to_numpy(), f.var(axis=0, ddof=1), tol * 8)
        close(f"C ABI [{label}] std_axis0", par.reduce_axis0_sharded(csf, Red.STD, n0, 0.0).to_numpy(), f.std(axis=0), tol * 8)
        cw = par.shard_axis0(Array.from_numpy(wide, cctx), rank, world)
        close(f"C ABI [{label}] var_axis0 wide", par.reduce_axis0_sharded(cw, Red.VAR, n0, 1.0).to_numpy(), wide.var(axis=0, ddof=1), tol * 8)
        # fewer rows than ranks: a rank with an EMPTY shard still takes part (min / max contribute the type's extreme)
        rows_t = max(1, world // 2)
        ti = rng.integers(-1000, 1000, size=(rows_t, 256), dtype=np.int64)
        tf = rng.standard_normal((rows_t, 256))
        sti = par.shard_axis0(Array.from_numpy(ti, cctx), rank, world)
        stf = par.shard_axis0(Array.from_numpy(tf, cctx), rank, world)
        for name, red, fn in (("min", Red.MIN, np.min), ("max", Red.MAX, np.max), ("sum", Red.SUM, np.sum)):
            ok(f"C ABI [{label}] {name}_axis0 i64, {rows_t} rows on {world} ranks",
               np.array_equal(par.reduce_axis0_sharded(sti, red, rows_t).to_numpy(), fn(ti, axis=0)))
            ok(f"C ABI [{label}] {name}() i64, {rows_t} rows on {world} ranks", par.reduce_all_sharded(sti, red) == fn(ti))
            if name != "sum":
                ok(f"C ABI [{label}] {name}_axis0 f64, {rows_t} rows on {world} ranks",
                   np.array_equal(par.reduce_axis0_sharded(stf, red, rows_t).to_numpy(), fn(tf, axis=0)))
                ok(f"C ABI [{label}] {name}() f64, {rows_t} rows on {world} ranks", par.reduce_all_sharded(stf, red) == fn(tf))
        # a kept axis: no exchange, outputs sharded like the input
        kept = Array.uninit((csf.shape[0],), np.float64, ctx=cctx)
        from ndarray_b200._ffi import check as _check, lib as _lib
        import ctypes as _C
        _check(_lib().ndb_reduce_axis_sharded(cctx.handle, Red.SUM, 1, _C.byref(csf.v), _C.byref(kept.v), 0.0, n0))
        close(f"C ABI [{label}] sum_axis1 (kept axis)", kept.to_numpy(), f[lo:hi].sum(axis=1), tol * 8)
        if gemm:
            for dt, tol_mm in ((np.float64, 1e-12), (np.float32, 1e-5)):
                A = rng.standard_normal((n0, 512)).astype(dt)
                B = rng.standard_normal((512, 300)).astype(dt)
                C0 = rng.standard_normal((n0, 300)).astype(dt)
                ref = (-2.3 * (A.astype(np.float64) @ B.astype(np.float64)) + np.pi * C0)[lo:hi]
                dA = par.shard_axis0(Array.from_numpy(A, cctx), rank, world)
                dC = par.shard_axis0(Array.from_numpy(C0, cctx), rank, world)
                dB = Array.from_numpy(B if rank == 0 else np.zeros_like(B), cctx)  # only rank 0 holds B
                par.gemm_rowsharded(-2.3, dA, dB, np.pi, dC, src=0)
                err = np.max(np.abs(dC.to_numpy() - ref)) / np.max(np.abs(ref)) if hi > lo else 0.0
                ok(f"C ABI [{label}] ndb_gemm_rowsharded {np.dtype(dt).name}", err <= tol_mm * np.sqrt(512), f"rel err {err:.3e}")
        cctx.sync()
        dist.barrier()
        comm.close()
    x = rng.standard_normal(n0 * 300)
    lo1, hi1 = par.shard_bounds(x.size, rank, world)
    dx = Array.from_numpy(x[lo1:hi1], ctx)
    close("global_dot", par.global_dot(dx, dx, dist), x @ x, 1e-12 * np.sqrt(x.size) * 8 * (x @ x))
    ctx.sync()
    dist.barrier()
    if own_peers:
        peers.close()
    everyone = [None] * world
    dist.all_gather_object(everyone, fails)
    total = [None] * world
    dist.all_gather_object(total, count[0])
    return sum(total), [m for fl in everyone for m in fl]


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["NDB_DEVICE"] = str(local)
    import ndarray_b200 as nd
    ctx = nd.Context(local)
    n, fails = run_checks(ctx, dist, rank, world)
    if rank == 0:
        for m in fails:
            print("FAIL", m)
        print(f"dist_check {'ok' if not fails else 'FAILED'} on {world} GPUs: {n} checks, {len(fails)} failures", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
