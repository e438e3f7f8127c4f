"""Outer-axis sharding across GPUs: one process per GPU, `torch.distributed` (NCCL) for the exchange steps.

ndarray's only parallelism is rayon's recursive `Zip::split` along the outermost axis
(src/zip/mod.rs:811-815,883-895; src/parallel/par.rs:270-297).  Across GPUs the same partition is made once,
with `split_at(Axis(0), …)` (src/impl_raw_views.rs:113-135): elementwise Zips then need no communication;
reductions whose reduced axis is the sharded one finish with one small collective.
torch is imported lazily and only here (plumbing for the collectives); all arithmetic stays in libndb200.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import View, check, lib
from .array import Array, Red, Slice


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of an axis of length n owned by `rank` (balanced, contiguous, in rank order)."""
    return n * rank // world, n * (rank + 1) // world


def shard_axis0(a: Array, rank: int, world: int) -> Array:
    """This rank's slab of `a` along Axis(0): two split_at calls, zero-copy.  Operands that are broadcast along
    axis 0 (stride 0) are replicated, not split, exactly as their stride says."""
    lo, hi = shard_bounds(a.shape[0], rank, world)
    _, right = a.split_at(0, lo)
    mine, _ = right.split_at(0, hi - lo)
    return mine


def _torch_view(arr: Array):
    """A torch tensor aliasing a dense device Array (for NCCL)."""
    import torch
    assert arr.is_standard_layout() or arr.ndim == 0
    n = arr.len()
    dt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
          np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[arr.dtype]

    class _Cai:  # __cuda_array_interface__ shim
        pass

    obj = _Cai()
    obj.__cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": arr.dtype.str, "data": (arr.v.ptr, False), "version": 3}
    t = torch.as_tensor(obj, device=torch.device("cuda", arr.ctx.device))
    assert t.dtype == dt
    return t


def _on_ctx_stream(arr: Array):
    import torch
    return torch.cuda.stream(torch.cuda.ExternalStream(arr.ctx.stream_ptr(), device=torch.device("cuda", arr.ctx.device)))


class PeerGroup:
    """Peer windows over NVLink for one context per rank (include/ndb200.h, csrc/peer.cu): every rank allocates a window
    in its own HBM, the 64-byte CUDA IPC handles are exchanged once through torch.distributed (host side, any backend),
    and from then on the small exchanges that end the sharded reductions are single kernel launches with no NCCL call.
    All ranks must issue the same sequence of exchanges."""

    def __init__(self, ctx, dist, slot_bytes: int = 1 << 20):
        self.ctx, self.rank, self.world = ctx, dist.get_rank(), dist.get_world_size()
        self._dist = dist
        # Collective-safe construction: a failure on ONE rank (IPC not permitted, out of memory) must not leave the others
        # waiting in a collective, so every step's outcome is exchanged and all ranks raise together.
        handle = (C.c_ubyte * 64)()
        err = None
        try:
            check(lib().ndb_peer_window_create(ctx.handle, self.rank, self.world, slot_bytes, handle))
        except Exception as e:  # noqa: BLE001
            err = repr(e)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (err, bytes(handle)))
        if any(g[0] for g in gathered):
            lib().ndb_peer_window_destroy(ctx.handle)
            raise RuntimeError(f"peer window creation failed on some rank: {[g[0] for g in gathered]}")
        blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(g[1] for g in gathered))
        try:
            check(lib().ndb_peer_window_attach(ctx.handle, blob))
        except Exception as e:  # noqa: BLE001
            err = repr(e)
        errs = [None] * self.world
        dist.all_gather_object(errs, err)  # doubles as the barrier: nobody stores into a window before all have mapped it
        if any(errs):
            lib().ndb_peer_window_destroy(ctx.handle)
            raise RuntimeError(f"peer window attach failed on some rank: {errs}")
        self.slot_bytes = slot_bytes

    def allreduce_(self, arr: Array, redop: int = Red.SUM) -> Array:
        """In-place all-reduce (rank order) of a dense device array of at most slot_bytes."""
        from .array import DTYPE_CODE as _DTYPE_CODE
        assert arr.is_standard_layout() or arr.ndim == 0
        check(lib().ndb_peer_allreduce(self.ctx.handle, redop, _DTYPE_CODE[arr.dtype], C.c_void_p(arr.v.ptr), C.c_void_p(arr.v.ptr),
                                       max(arr.len(), 1)))
        return arr

    def allgather(self, arr: Array) -> Array:
        """[world, *arr.shape] with every rank's dense array, in rank order."""
        from .array import DTYPE_CODE as _DTYPE_CODE
        assert arr.is_standard_layout() or arr.ndim == 0
        out = Array.uninit((self.world,) + tuple(arr.shape), arr.dtype, ctx=self.ctx)
        check(lib().ndb_peer_allgather(self.ctx.handle, _DTYPE_CODE[arr.dtype], C.c_void_p(arr.v.ptr), C.c_void_p(out.v.ptr),
                                       max(arr.len(), 1)))
        return out

    def close(self):
        check(lib().ndb_peer_window_destroy(self.ctx.handle))


class Comm:
    """The library-owned NCCL communicator of one context (csrc/dist.cu): what a non-Python host would set up with
    ndb_comm_unique_id / ndb_comm_init.  Here the 128-byte id travels through torch.distributed's object broadcast (host side,
    any backend); after that the sharded entry points (ndb_gemm_rowsharded, ndb_reduce_axis_sharded, ndb_reduce_all_sharded)
    involve no Python collectives at all."""

    def __init__(self, ctx, dist):
        self.ctx, self.rank, self.world = ctx, dist.get_rank(), dist.get_world_size()
        uid = (C.c_ubyte * 128)()
        err = None
        if self.rank == 0:
            try:
                check(lib().ndb_comm_unique_id(uid))
            except Exception as e:  # noqa: BLE001
                err = repr(e)
        box = [(err, bytes(uid))]
        dist.broadcast_object_list(box, src=0)
        if box[0][0]:
            raise RuntimeError(f"ndb_comm_unique_id failed on rank 0: {box[0][0]}")
        blob = (C.c_ubyte * 128).from_buffer_copy(box[0][1])
        try:
            check(lib().ndb_comm_init(ctx.handle, self.rank, self.world, blob))
        except Exception as e:  # noqa: BLE001
            err = repr(e)
        errs = [None] * self.world
        dist.all_gather_object(errs, err)
        if any(errs):
            lib().ndb_comm_destroy(ctx.handle)
            raise RuntimeError(f"ndb_comm_init failed on some rank: {errs}")

    def close(self):
        check(lib().ndb_comm_destroy(self.ctx.handle))


def gemm_rowsharded(alpha, a_local: Array, b: Array, beta, c_local: Array, src: int = 0) -> None:
    """ndb_gemm_rowsharded: B (valid on rank `src`) is broadcast by the library's own communicator, then the local GEMM."""
    al, be = np.asarray(alpha, dtype=a_local.dtype), np.asarray(beta, dtype=a_local.dtype)
    check(lib().ndb_gemm_rowsharded(a_local.ctx.handle, C.c_void_p(al.ctypes.data), C.byref(a_local.v), C.byref(b.v),
                                    C.c_void_p(be.ctypes.data), C.byref(c_local.v), int(src)))


def reduce_axis0_sharded(local: Array, redop: int, n_global: int, ddof: float = 0.0) -> Array:
    """ndb_reduce_axis_sharded along the sharded axis: fused over peer windows when possible, NCCL otherwise - decided
    inside the library, identically on every rank."""
    out = Array.uninit(local.shape[1:], local.dtype, ctx=local.ctx)
    check(lib().ndb_reduce_axis_sharded(local.ctx.handle, redop, 0, C.byref(local.v), C.byref(out.v), float(ddof), int(n_global)))
    return out


def reduce_all_sharded(local: Array, redop: int = Red.SUM):
    out = np.zeros((), dtype=local.dtype)
    check(lib().ndb_reduce_all_sharded(local.ctx.handle, redop, C.byref(local.v), C.c_void_p(out.ctypes.data)))
    return out[()]


def _fused_axis0(local: Array, redop: int, ddof: float, n_global: int, peers: "PeerGroup"):
    """ndb_reduce_axis_peer along Axis(0): the whole sharded reduction (local column reduction, exchange of the per-tile
    states through peer memory, rank-ordered merge, finalisation with the global length) in ONE cooperative launch.
    Returns None when the library rules the layout out.  The library's rule uses only quantities that are the same on
    every rank for shards of one global array (kept shape, strides, dtype, op, n_global, window size, SM count): a
    rank-local accident - a shard that starts off a 16-byte boundary, holds a single row or is empty - changes how that
    rank loads its rows, never whether it takes part, so no cross-rank vote is needed and no rank can be left waiting."""
    from . import _ffi
    out = Array.uninit(local.shape[1:], local.dtype, ctx=local.ctx)
    st = lib().ndb_reduce_axis_peer(local.ctx.handle, redop, 0, C.byref(local.v), C.byref(out.v), float(ddof), int(n_global))
    if st == _ffi.ERR_UNSUPPORTED:
        return None
    check(st)
    return out


def _multi(dist) -> bool:
    return dist is not None and dist.is_initialized() and dist.get_world_size() > 1


def global_sum(local: Array, dist=None, peers: PeerGroup | None = None):
    """sum() over all shards.  With a PeerGroup the kernel that finishes this rank's reduction tree also exchanges the
    scalar through peer memory (ndb_reduce_all_peer: no extra launch, no NCCL); otherwise an NCCL all-reduce of ONE
    element follows on the same stream."""
    out = Array.zeros((), local.dtype, ctx=local.ctx)
    if peers is not None and peers.world > 1:
        check(lib().ndb_reduce_all_peer(local.ctx.handle, Red.SUM, C.byref(local.v), C.c_void_p(out.v.ptr)))
        return out.to_numpy()[()]
    check(lib().ndb_reduce_all_dev(local.ctx.handle, Red.SUM, C.byref(local.v), C.c_void_p(out.v.ptr)))
    if _multi(dist):
        with _on_ctx_stream(out):
            dist.all_reduce(_torch_view(out))
    return out.to_numpy()[()]


def sum_axis0(local: Array, dist=None, peers: PeerGroup | None = None, n_global: int | None = None) -> Array:
    """sum_axis(Axis(0)) when axis 0 is the sharded axis: local partial vector, then all-reduce(sum) — one peer-memory
    exchange kernel with a PeerGroup, NCCL otherwise.  With the GLOBAL axis length the whole thing is one fused launch;
    whether that path is taken is decided from n_global and the (rank-invariant) layout only, never from this rank's own
    row count, so ranks cannot diverge."""
    if peers is not None and peers.world > 1 and n_global is not None and n_global > 0:
        fused = _fused_axis0(local, Red.SUM, 0.0, n_global, peers)
        if fused is not None:
            return fused
    part = local.sum_axis(0)
    if peers is not None and peers.world > 1 and part.len() * part.dtype.itemsize <= peers.slot_bytes:
        return peers.allreduce_(part)
    if _multi(dist):
        with _on_ctx_stream(part):
            dist.all_reduce(_torch_view(part))
    return part


def mean_axis0(local: Array, n_global: int, dist=None, peers: PeerGroup | None = None) -> Array:
    """mean_axis(Axis(0)): divide by the GLOBAL length after the all-reduce (impl_numeric.rs:342-355)."""
    if peers is not None and peers.world > 1 and n_global > 0 and local.dtype.kind == "f":
        fused = _fused_axis0(local, Red.MEAN, 0.0, n_global, peers)
        if fused is not None:
            return fused
    s = sum_axis0(local, dist, peers, n_global)
    s /= local.dtype.type(n_global)
    return s


def var_axis0(local: Array, ddof: float, dist=None, want_std: bool = False, peers: PeerGroup | None = None,
              n_global: int | None = None) -> Array:
    """var_axis(Axis(0), ddof): per-shard Welford state (n, mean, M2), all-gather, Chan merge in rank order.  With a
    PeerGroup and the global axis length the whole thing is one launch (ndb_reduce_axis_peer)."""
    import torch
    shape = local.shape[1:]
    if peers is not None and peers.world > 1 and n_global is not None and n_global > 0:
        fused = _fused_axis0(local, Red.STD if want_std else Red.VAR, ddof, n_global, peers)
        if fused is not None:
            return fused
    if peers is not None and peers.world > 1 and max(int(np.prod(shape)), 1) * local.dtype.itemsize <= peers.slot_bytes:
        mean, m2 = Array.uninit(shape, local.dtype, ctx=local.ctx), Array.uninit(shape, local.dtype, ctx=local.ctx)
        check(lib().ndb_welford_axis(local.ctx.handle, 0, C.byref(local.v), C.byref(mean.v), C.byref(m2.v)))
        gm, g2 = peers.allgather(mean), peers.allgather(m2)
        cnt = peers.allgather(Array.from_numpy(np.array(local.shape[0], dtype=np.int64), local.ctx))
        counts = [int(x) for x in cnt.to_numpy().reshape(-1)]
        means = [gm.index_axis(0, r) for r in range(peers.world)]
        m2s = [g2.index_axis(0, r) for r in range(peers.world)]
        out = Array.uninit(shape, local.dtype, ctx=local.ctx)
        mv = (View * len(means))(*[m.v for m in means])
        sv = (View * len(m2s))(*[m.v for m in m2s])
        cv = (C.c_int64 * len(counts))(*counts)
        check(lib().ndb_welford_merge(local.ctx.handle, len(counts), cv, mv, sv, float(ddof), int(want_std), C.byref(out.v)))
        return out
    mean, m2 = Array.uninit(shape, local.dtype, ctx=local.ctx), Array.uninit(shape, local.dtype, ctx=local.ctx)
    check(lib().ndb_welford_axis(local.ctx.handle, 0, C.byref(local.v), C.byref(mean.v), C.byref(m2.v)))
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    counts = [local.shape[0]]
    means, m2s = [mean], [m2]
    if world > 1:
        n_out = max(mean.len(), 1)
        gm = torch.empty((world, n_out), dtype=_torch_view(mean).dtype, device=_torch_view(mean).device)
        g2 = torch.empty_like(gm)
        cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=gm.device)
        gc = torch.empty((world,), dtype=torch.int64, device=gm.device)
        with _on_ctx_stream(mean):
            dist.all_gather_into_tensor(gm, _torch_view(mean))
            dist.all_gather_into_tensor(g2, _torch_view(m2))
            dist.all_gather_into_tensor(gc, cnt)
        local.ctx.sync()  # the gathers ran on the context stream; the host reads the counts next
        counts = [int(x) for x in gc.tolist()]
        means = [Array.borrow(gm[r].data_ptr(), shape, local.dtype, ctx=local.ctx, owner=gm) for r in range(world)]
        m2s = [Array.borrow(g2[r].data_ptr(), shape, local.dtype, ctx=local.ctx, owner=g2) for r in range(world)]
    out = Array.uninit(shape, local.dtype, ctx=local.ctx)
    mv = (View * len(means))(*[m.v for m in means])
    sv = (View * len(m2s))(*[m.v for m in m2s])
    cv = (C.c_int64 * len(counts))(*counts)
    check(lib().ndb_welford_merge(local.ctx.handle, len(counts), cv, mv, sv, float(ddof), int(want_std), C.byref(out.v)))
    return out


def global_dot(x_local: Array, y_local: Array, dist=None):
    """1-D dot over sharded vectors: per-GPU dot (src/linalg/impl_linalg.rs:74-133), then all-reduce of one element."""
    import torch
    part = x_local.dot(y_local)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([part], dtype=_torch_view(x_local).dtype, device=torch.device("cuda", x_local.ctx.device))
        dist.all_reduce(t)
        return t.cpu().numpy()[0]
    return part


def general_mat_mul_rowsharded(alpha, a_local: Array, b: Array, beta, c_local: Array, dist=None, src: int = 0,
                               k_chunks: int = 1) -> None:
    """general_mat_mul with A and C sharded along Axis(0) and B replicated: `b` is a C-order [k, n]
    buffer on every rank whose contents are valid on rank `src`; it is broadcast over NCCL and each rank multiplies its
    own rows.  With k_chunks > 1 the broadcast is cut into row blocks of B on a second stream and the GEMM on block j
    (C += alpha * A[:, kj] * B[kj, :]) overlaps the transfer of block j + 1, so the link time hides behind the tensor
    pipe; the chunked sum differs from the one-shot result only in rounding order (GEMM parity is to tolerance)."""
    from .linalg import general_mat_mul
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    k = b.shape[0]
    if world == 1:
        general_mat_mul(alpha, a_local, b, beta, c_local)
        return
    import torch
    assert b.is_standard_layout(), "the replicated operand is broadcast as one dense C-order buffer"
    dev = torch.device("cuda", b.ctx.device)
    tb = _torch_view(b).view(k, -1)
    compute = torch.cuda.ExternalStream(b.ctx.stream_ptr(), device=dev)
    k_chunks = max(1, min(int(k_chunks), k))
    if k_chunks == 1:
        with torch.cuda.stream(compute):
            dist.broadcast(tb, src)
        general_mat_mul(alpha, a_local, b, beta, c_local)
        return
    comm = torch.cuda.Stream(device=dev)
    ready = torch.cuda.Event()
    ready.record(compute)        # B was produced on the context stream
    comm.wait_event(ready)
    bounds = [k * j // k_chunks for j in range(k_chunks + 1)]
    for j in range(k_chunks):
        lo, hi = bounds[j], bounds[j + 1]
        if hi == lo:
            continue
        with torch.cuda.stream(comm):
            dist.broadcast(tb[lo:hi], src)
            landed = torch.cuda.Event()
            landed.record(comm)
        compute.wait_event(landed)
        a_blk = a_local.slice_axis(1, Slice(lo, hi, 1))
        b_blk = b.slice_axis(0, Slice(lo, hi, 1))
        general_mat_mul(alpha, a_blk, b_blk, beta if j == 0 else 1.0, c_local)
    done = torch.cuda.Event()
    done.record(compute)
    comm.wait_event(done)        # the comm stream (and its tensors) must outlive the last GEMM
