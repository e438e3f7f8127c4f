"""world_size-2 `gloo` tests (CPU) of the multi-GPU protocol:

* the partition: `shard_bounds` + `ndb_view_split_at` through the C ABI give disjoint slabs that tile the array;
* elementwise Zip on shards needs no exchange: per-shard oracle results concatenate to the unsharded result bit-exactly;
* sum / sum_axis(Axis(0)): per-shard partials + all-reduce reproduce the unsharded oracle (exactly for integers);
* var_axis(Axis(0)): per-shard (n, mean, M2) + all-gather + Chan merge in rank order reproduces Welford;
* general_mat_mul: A and C row-sharded, B broadcast from rank 0 (one shot, and accumulated in row blocks of B).
Local compute uses the oracle here (no GPU); on the GPU box the same protocol runs in ndarray_b200/parallel.py."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from ndarray_b200 import _ffi
from ndarray_b200._ffi import View
from ndarray_b200.parallel import shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def chan_merge(states):
    """(n, mean, M2) merged in order with the formula of csrc/reduce.cu (Reducer<T, R_WELFORD>::merge)."""
    n, mean, m2 = states[0]
    mean, m2 = mean.copy(), m2.copy()
    for nb, mb, m2b in states[1:]:
        if nb == 0:
            continue
        if n == 0:
            n, mean, m2 = nb, mb.copy(), m2b.copy()
            continue
        tot = n + nb
        delta = mb - mean
        w = nb / tot
        mean = mean + delta * w
        m2 = m2 + m2b + delta * delta * (n * w)
        n = tot
    return n, mean, m2


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same data on every rank; each rank takes its slab
        a = rng.integers(-2 ** 62, 2 ** 62, size=(37, 5, 6), dtype=np.int64)
        b = rng.integers(-2 ** 62, 2 ** 62, size=(37, 5, 6), dtype=np.int64)
        f = rng.standard_normal((37, 11))
        lo, hi = shard_bounds(37, rank, world)
        # elementwise: no exchange
        out = np.zeros((hi - lo, 5, 6), np.int64)
        assert orc.zip_(2, out, [a[lo:hi], b[lo:hi]]) == 0
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        full = np.zeros_like(a)
        assert orc.zip_(2, full, [a, b]) == 0
        assert np.array_equal(np.concatenate(gathered, axis=0), full)
        # global sum: scalar all-reduce (wrapping i64 is associative: exact)
        part = torch.tensor([int(orc.reduce_all(0, a[lo:hi]))], dtype=torch.int64)
        dist.all_reduce(part)
        assert int(part.item()) == int(orc.reduce_all(0, a))
        # sum_axis(Axis(0)): vector all-reduce
        st, p0 = orc.reduce_axis(0, a[lo:hi], 0)
        t = torch.from_numpy(p0.copy())
        dist.all_reduce(t)
        assert np.array_equal(t.numpy(), orc.reduce_axis(0, a, 0)[1])
        # var_axis(Axis(0)): Welford shards + Chan merge in rank order
        st, mean_l = orc.reduce_axis(4, f[lo:hi], 0)
        st, var_l = orc.reduce_axis(5, f[lo:hi], 0, 0.0)
        state = (hi - lo, mean_l, var_l * (hi - lo))
        states = [None] * world
        dist.all_gather_object(states, state)
        n, mean, m2 = chan_merge(states)
        want = orc.reduce_axis(5, f, 0, 1.0)[1]
        assert n == 37
        assert np.max(np.abs(m2 / (n - 1.0) - want)) <= 1e-12 * np.sqrt(37) * np.abs(want).max() * 8
        # general_mat_mul: A and C row-sharded, B valid on rank 0 only and broadcast; slabs concatenate to the full product
        A = rng.standard_normal((37, 16))
        B = rng.standard_normal((16, 9))
        C0 = rng.standard_normal((37, 9))
        tb = torch.from_numpy(B.copy() if rank == 0 else np.zeros_like(B))
        dist.broadcast(tb, 0)
        c_loc = C0[lo:hi].copy()
        assert orc.gemm(-2.3, A[lo:hi], tb.numpy(), np.pi, c_loc) == 0
        slabs = [None] * world
        dist.all_gather_object(slabs, c_loc)
        c_full = C0.copy()
        assert orc.gemm(-2.3, A, B, np.pi, c_full) == 0
        assert np.array_equal(np.concatenate(slabs, axis=0), c_full)  # row sharding does not change any row's arithmetic
        # the pipelined form adds B's row blocks one at a time: same result within rounding
        c_chunk = C0[lo:hi].copy()
        for j, (k0, k1) in enumerate(((0, 4), (4, 8), (8, 12), (12, 16))):
            assert orc.gemm(-2.3, A[lo:hi, k0:k1], B[k0:k1], np.pi if j == 0 else 1.0, c_chunk) == 0
        assert np.max(np.abs(c_chunk - c_full[lo:hi])) <= 1e-12 * np.sqrt(16) * np.abs(c_full).max()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


@pytest.mark.parametrize("n,world", [(37, 2), (8, 8), (5, 8), (65536, 4), (0, 2)])
def test_shard_bounds_tile_the_axis(n, world):
    b = [shard_bounds(n, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == n
    for (l0, h0), (l1, h1) in zip(b, b[1:]):
        assert h0 == l1 and l0 <= h0
    sizes = [h - l for l, h in b]
    assert max(sizes) - min(sizes) <= 1


def test_split_at_slabs_through_the_abi():
    """The slabs ndb_view_split_at produces for each rank are disjoint and adjacent, also for a reversed axis."""
    L = _ffi.lib()
    for stride0, base in ((30, 4096), (-30, 4096 + 36 * 30 * 8)):
        v = View()
        v.ptr, v.ndim, v.dtype = base, 3, 1
        for i, (d, s) in enumerate(zip((37, 5, 6), (stride0, 6, 1))):
            v.dim[i], v.stride[i] = d, s
        world = 4
        ptrs = []
        for r in range(world):
            lo, hi = shard_bounds(37, r, world)
            l, rest, mine, tail = View(), View(), View(), View()
            assert L.ndb_view_split_at(C.byref(v), 0, lo, C.byref(l), C.byref(rest)) == 0
            assert L.ndb_view_split_at(C.byref(rest), 0, hi - lo, C.byref(mine), C.byref(tail)) == 0
            assert mine.dim[0] == hi - lo and mine.stride[0] == stride0
            assert mine.ptr == base + lo * stride0 * 8
            ptrs.append((mine.ptr, mine.dim[0]))
        assert sum(d for _, d in ptrs) == 37
