// Reductions for sm_100a: sum / product / min / max over a whole view, and along one axis
// (sum_axis, product_axis, mean_axis, var_axis, std_axis, fold_axis with an associative closure).
//
// ndarray (src/numeric/impl_numeric.rs:32-47,268-283,342-355,401-425; src/impl_methods.rs:3130-3141) runs
// these on one core: an 8-accumulator unrolled fold per contiguous lane, or a sequential pass over
// subviews.  Here every reduction is a tree: registers -> warp shuffles -> shared memory -> (when the
// reduced axis is split across blocks) a second, ordered pass over per-block partial states.  All
// shapes of tree are fixed by the problem size alone, so results are run-to-run deterministic.
//
// Kernel classes
//   red_lanes   the reduced axis is the densest axis of the input (|stride| smallest): a warp or a
//               block walks each lane with coalesced (optionally 128-bit) loads.
//   red_cols    a KEPT axis is unit-stride: threads own output columns and stride down the reduced
//               axis; the reduced axis is additionally split over blockIdx.y so that 16384 columns
//               still fill 148 SMs; partial states are merged in split order by red_finish.
//   red_generic one thread per output, any strides (correctness net).
//   red_all_*   whole-view reduction: grid-stride partials + red_finish.
// var_axis/std_axis use a fused single-pass Welford per thread and Chan's pairwise merge up the tree.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "zip_kernels.cuh"  // w_add & friends, Vec128
#include "peer.cuh"

// tuning knobs (compile-time so that benchmarks/build_variant.sh can sweep them)
#ifndef NDB_RED_LANES_U
#define NDB_RED_LANES_U 8
#endif
#ifndef NDB_RED_COLS_MINB_W
#define NDB_RED_COLS_MINB_W 3
#endif
#ifndef NDB_RED_COLS_MINB
#define NDB_RED_COLS_MINB 3
#endif
#ifndef NDB_RED_COLS_U
#define NDB_RED_COLS_U 8
#endif
#ifndef NDB_ZIP_FOLD_U
#define NDB_ZIP_FOLD_U 2
#endif
#ifndef NDB_RED_COLS_WU
#define NDB_RED_COLS_WU 8
#endif
#include "zip_plan.h"

namespace ndb {

enum { R_SUM = 0, R_PROD = 1, R_MIN = 2, R_MAX = 3, R_WELFORD = 4 };
enum { FIN_NONE = 0, FIN_MEAN = 1, FIN_VAR = 2, FIN_STD = 3, FIN_WELFORD_RAW = 4 };

template <typename T>
struct WelfordState {
    T n, mean, m2;
};

template <typename T, int R>
struct Reducer {
    using State = T;
    __device__ __forceinline__ static State init()
    {
        if constexpr (R == R_SUM) return T(0);
        else if constexpr (R == R_PROD) return T(1);
        else return T(0);  // MIN/MAX seed from the first element, see `seeded`
    }
    __device__ __forceinline__ static void add(State &s, T x)
    {
        if constexpr (R == R_SUM) s = w_add(s, x);
        else if constexpr (R == R_PROD) s = w_mul(s, x);
        else if constexpr (R == R_MIN) s = w_min(s, x);
        else s = w_max(s, x);
    }
    __device__ __forceinline__ static void merge(State &s, const State &o) { add(s, o); }
    __device__ __forceinline__ static State shfl_down(const State &s, int d) { return __shfl_down_sync(0xffffffffu, s, d); }
};

template <typename T>
struct Reducer<T, R_WELFORD> {
    using State = WelfordState<T>;
    __device__ __forceinline__ static State init() { return State{T(0), T(0), T(0)}; }
    // one step of ndarray's loop (impl_numeric.rs:419-421), fused multiply-add included
    __device__ __forceinline__ static void add(State &s, T x)
    {
        s.n += T(1);
        const T delta = x - s.mean;
        s.mean = s.mean + delta / s.n;
        if constexpr (std::is_same<T, float>::value) s.m2 = fmaf(x - s.mean, delta, s.m2);
        else s.m2 = fma(x - s.mean, delta, s.m2);
    }
    // Chan et al. pairwise combination
    __device__ __forceinline__ static void merge(State &s, const State &o)
    {
        if (o.n == T(0)) return;
        if (s.n == T(0)) {
            s = o;
            return;
        }
        const T n = s.n + o.n;
        const T delta = o.mean - s.mean;
        const T w = o.n / n;
        s.mean = s.mean + delta * w;
        s.m2 = s.m2 + o.m2 + delta * delta * (s.n * w);
        s.n = n;
    }
    __device__ __forceinline__ static State shfl_down(const State &s, int d)
    {
        State r;
        r.n = __shfl_down_sync(0xffffffffu, s.n, d);
        r.mean = __shfl_down_sync(0xffffffffu, s.mean, d);
        r.m2 = __shfl_down_sync(0xffffffffu, s.m2, d);
        return r;
    }
};

// MIN/MAX have no identity that is safe for every dtype and NaN rule, so states carry "seeded by the
// first element": every thread that owns at least one element starts from it; threads that own none
// are skipped in merges via a validity flag folded into the tree.
template <typename T, int R>
struct Acc {
    using Rd = Reducer<T, R>;
    typename Rd::State s;
    bool valid;
    __device__ __forceinline__ void reset()
    {
        s = Rd::init();
        valid = (R != R_MIN && R != R_MAX);
    }
    __device__ __forceinline__ void add(T x)
    {
        if constexpr (R == R_MIN || R == R_MAX) {
            if (!valid) {
                s = x;
                valid = true;
                return;
            }
        }
        Rd::add(s, x);
    }
    // K elements at once.  Welford: two-pass statistics of the batch (exact mean of K values, squared
    // deviations from it), then ONE Chan merge - one division per K elements instead of one per element, and a
    // smaller rounding error than K sequential updates.  Other reducers: K sequential adds.
    template <int K>
    __device__ __forceinline__ void add_batch(const T (&x)[K])
    {
        if constexpr (R == R_WELFORD) {
            T sum = x[0];
#pragma unroll
            for (int i = 1; i < K; ++i) sum += x[i];
            const T bmean = sum * (T(1) / T(K));
            T bm2 = T(0);
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const T d = x[i] - bmean;
                bm2 += d * d;
            }
            const T n = s.n + T(K);
            const T delta = bmean - s.mean;
            const T w = T(K) / n;
            s.mean = s.mean + delta * w;
            s.m2 = s.m2 + bm2 + delta * delta * (s.n * w);
            s.n = n;
        } else {
            // pairwise tree over the batch, then ONE add into the running state: the sequential chain a thread carries is
            // K times shorter, so is its rounding-error random walk (2^32-element f32 sum: ~0.16 -> ~0.04 absolute), and the
            // K values combine with log2(K) dependent operations instead of K
            T t[K];
#pragma unroll
            for (int i = 0; i < K; ++i) t[i] = x[i];
#pragma unroll
            for (int w = K; w > 1; w = (w + 1) / 2) {
#pragma unroll
                for (int i = 0; i < w / 2; ++i) {
                    typename Rd::State m = t[2 * i];
                    Rd::add(m, t[2 * i + 1]);
                    t[i] = m;
                }
                if (w & 1) t[w / 2] = t[w - 1];
            }
            add(t[0]);
        }
    }
    __device__ __forceinline__ void merge(const Acc &o)
    {
        if constexpr (R == R_MIN || R == R_MAX) {
            if (!o.valid) return;
            if (!valid) {
                s = o.s;
                valid = true;
                return;
            }
        }
        Rd::merge(s, o.s);
    }
    __device__ __forceinline__ Acc shfl_down(int d) const
    {
        Acc r;
        r.s = Rd::shfl_down(s, d);
        r.valid = __shfl_down_sync(0xffffffffu, (int)valid, d) != 0;
        return r;
    }
};

template <typename T, int R>
__device__ __forceinline__ Acc<T, R> warp_reduce(Acc<T, R> a)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Acc<T, R> o = a.shfl_down(d);
        a.merge(o);
    }
    return a;  // lane 0 holds the result; lanes merge in a fixed order
}

// Block-wide: every warp reduces, warp leaders park in shared memory, warp 0 finishes. Result valid in thread 0.
template <typename T, int R>
__device__ __forceinline__ Acc<T, R> block_reduce(Acc<T, R> a, Acc<T, R> *smem /* >= 32 entries */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    a = warp_reduce(a);
    if (lane == 0) smem[warp] = a;
    __syncthreads();
    if (warp == 0) {
        Acc<T, R> b;
        b.reset();
        if (lane < nwarps) b = smem[lane];
        b = warp_reduce(b);
        a = b;
    }
    __syncthreads();
    return a;
}

template <typename T, int R, int FIN>
__device__ __forceinline__ T finalize(const Acc<T, R> &a, T nT, T dof)
{
    if constexpr (R == R_WELFORD) {
        T v = a.s.m2 / dof;
        if constexpr (FIN == FIN_STD) {
            if constexpr (std::is_same<T, float>::value) v = sqrtf(v);
            else v = sqrt(v);
        }
        return v;
    } else if constexpr (FIN == FIN_MEAN) {
        int dummy = 0;
        return w_div(a.s, nT, &dummy);
    } else {
        return a.s;
    }
}

// ---- kept-axes geometry ----------------------------------------------------------------------
struct Kept {
    int n;  // collapsed kept rank (>= 1)
    int64_t dim[NDB_MAX_NDIM];
    int64_t in_stride[NDB_MAX_NDIM];
    int64_t out_stride[NDB_MAX_NDIM];
    int64_t out2_stride[NDB_MAX_NDIM];
};

struct RedArgs {
    const char *in;
    char *out;
    char *out2;  // FIN_WELFORD_RAW: second output (m2); `out` receives the mean
    Kept k;
    int64_t n_out;  // product of kept dims
    int64_t n;      // length of the reduced axis
    int64_t s;      // its stride, elements
    double nT;      // n as the element type
    double dof;     // n - ddof in the element type
    int splits;     // red_cols: blocks along the reduced axis
    int lane_warps; // red_lanes<WARP>: warps sharing one lane (1, 2, 4 or 8)
    void *partial;  // [splits][n_out] states
    unsigned *tickets;  // red_cols: one zeroed counter per column tile; non-null = the last split to finish merges (one launch)
    int peer_on;                      // red_cols<PEER>: exchange every tile's state with the other ranks before finalizing
    int peer_scalar;                  // red_cols<PEER>: this rank's base pointer is not 16-byte aligned -> same tile geometry, scalar loads
    unsigned long long *tile_epoch;   // exchanges completed so far per tile counter (device, this rank)
    PeerArgs peer;
};

__device__ __forceinline__ void kept_offsets(const Kept &k, int64_t lin, int64_t &in_off, int64_t &out_off,
                                             int64_t &out2_off)
{
    if (k.n == 1) {  // one kept axis (every 2-d input): no 64-bit division in front of the first load
        in_off = lin * k.in_stride[0];
        out_off = lin * k.out_stride[0];
        out2_off = lin * k.out2_stride[0];
        return;
    }
    in_off = out_off = out2_off = 0;
    for (int d = k.n - 1; d >= 0; --d) {
        const int64_t q = lin / k.dim[d];
        const int64_t i = lin - q * k.dim[d];
        lin = q;
        in_off += i * k.in_stride[d];
        out_off += i * k.out_stride[d];
        out2_off += i * k.out2_stride[d];
    }
}

template <typename T, int R, int FIN>
__device__ __forceinline__ void store_result(const RedArgs &a, const Acc<T, R> &acc, int64_t out_off, int64_t out2_off)
{
    if constexpr (FIN == FIN_WELFORD_RAW) {
        reinterpret_cast<T *>(a.out)[out_off] = acc.s.mean;
        reinterpret_cast<T *>(a.out2)[out2_off] = acc.s.m2;
    } else {
        reinterpret_cast<T *>(a.out)[out_off] = finalize<T, R, FIN>(acc, (T)a.nT, (T)a.dof);
    }
}

// ---- lanes: reduced axis is the dense one -----------------------------------------------------
// WARP=true: one warp per output (short lanes); else one block per output.
template <typename T, int R, int FIN, bool WARP, bool VEC>
__global__ void __launch_bounds__(256) red_lanes(RedArgs a)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    __shared__ Acc<T, R> smem[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    // WARP: P = a.lane_warps warps share one lane (P = 1: a warp per lane), so a block covers wpb / P lanes; short
    // problems want every thread slot of the machine busy with a few loads each rather than few threads with many
    const int P = WARP ? a.lane_warps : 1;
    const int lanes_pb = wpb / P;
    const int64_t groups = WARP ? (a.n_out + lanes_pb - 1) / lanes_pb : a.n_out;
    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const int64_t o = WARP ? g * lanes_pb + warp / P : g;
        const bool live = o < a.n_out;
        int64_t in_off = 0, out_off = 0, out2_off = 0;
        if (live) kept_offsets(a.k, o, in_off, out_off, out2_off);
        const T *src = reinterpret_cast<const T *>(a.in) + in_off;
        const int tid = WARP ? (warp % P) * 32 + lane : threadIdx.x;
        const int nt = WARP ? 32 * P : blockDim.x;
        Acc<T, R> acc;
        acc.reset();
        if (live) {
            if constexpr (VEC) {
                // a.s == 1, lanes 16-byte aligned, n % N == 0 (checked on the host)
                const V *vsrc = reinterpret_cast<const V *>(src);
                const int64_t nv = a.n / N;
                int64_t i = tid;
                constexpr int U = (R == R_WELFORD) ? 4 : NDB_RED_LANES_U;  // vectors in flight per thread
                for (; i + (U - 1) * (int64_t)nt < nv; i += U * (int64_t)nt) {
                    V v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = vsrc[i + u * (int64_t)nt];
                    T x[U * N];
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int e = 0; e < N; ++e) x[u * N + e] = v[u].v[e];
                    acc.template add_batch<U * N>(x);
                }
                for (; i < nv; i += nt) {
                    V v0 = vsrc[i];
                    acc.template add_batch<N>(v0.v);
                }
            } else {
                int64_t i = tid;
                for (; i + 3 * (int64_t)nt < a.n; i += 4 * (int64_t)nt) {
                    const T x[4] = {src[i * a.s], src[(i + nt) * a.s], src[(i + 2 * nt) * a.s], src[(i + 3 * nt) * a.s]};
                    acc.template add_batch<4>(x);
                }
                for (; i < a.n; i += nt) acc.add(src[i * a.s]);
            }
        }
        if constexpr (WARP) {
            acc = warp_reduce(acc);
            if (P == 1) {
                if (live && lane == 0) store_result<T, R, FIN>(a, acc, out_off, out2_off);
            } else {  // the P warps of a lane meet in shared memory and merge in warp order
                if (lane == 0) smem[warp] = acc;
                __syncthreads();
                if (live && lane == 0 && warp % P == 0) {
                    for (int q = 1; q < P; ++q) acc.merge(smem[warp + q]);
                    store_result<T, R, FIN>(a, acc, out_off, out2_off);
                }
                __syncthreads();
            }
        } else {
            acc = block_reduce(acc, smem);
            if (threadIdx.x == 0) store_result<T, R, FIN>(a, acc, out_off, out2_off);
        }
    }
}

// a partial state written by another block: read through L2 (ld.global.cg), never from this SM's L1
template <typename A>
__device__ __forceinline__ A ldcg_state(const A *p)
{
    static_assert(sizeof(A) % 4 == 0, "partial states are whole words");
    A r;
    const unsigned *s = reinterpret_cast<const unsigned *>(p);
    unsigned *d = reinterpret_cast<unsigned *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(A) / 4); ++i) d[i] = __ldcg(s + i);
    return r;
}

// states p[0], p[pitch], ... merged in split order; loads are issued eight at a time so that the chain costs one L2
// round trip per eight splits instead of one per split (it sits on the kernel's critical path)
template <typename A>
__device__ __forceinline__ A merge_splits(const A *p, int splits, int64_t pitch)
{
    // batch size by state size: eight 24-byte Welford states in flight cost 48 registers and the third CTA per SM
    // (var_axis 6.6 -> 4.8 TB/s), so wide states go two at a time
    constexpr int B = sizeof(A) > 8 ? 2 : 8;
    A r = ldcg_state(p);
    int sidx = 1;
    for (; sidx + B <= splits; sidx += B) {
        A b[B];
#pragma unroll
        for (int u = 0; u < B; ++u) b[u] = ldcg_state(p + (int64_t)(sidx + u) * pitch);
#pragma unroll
        for (int u = 0; u < B; ++u) r.merge(b[u]);
    }
    for (; sidx < splits; ++sidx) r.merge(ldcg_state(p + (int64_t)sidx * pitch));
    return r;
}

// ---- cols: a kept axis is unit-stride ------------------------------------------------------------
// blockDim = (32, 8).  A block owns 32*W consecutive outputs along the last kept axis (in_stride 1) and
// rows [split*chunk, ...) of the reduced axis; thread (tx, ty) strides rows ty, ty+8, ...
template <typename T, int R, int FIN, int W, bool PEER = false>
__global__ void __launch_bounds__(256, R == R_WELFORD ? NDB_RED_COLS_MINB_W : NDB_RED_COLS_MINB) red_cols(RedArgs a)  // 3 CTAs/SM register budget (<= 85): f64 sum 6.52 -> 6.95 TB/s vs the default allocation
{
    pdl_enter();
    using V = Vec128<T>;
    static_assert(W == 1 || W == V::N, "W is 1 or one 128-bit vector");
    __shared__ Acc<T, R> smem[8][32 * W + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int last = a.k.n - 1;
    const int64_t ncol = a.k.dim[last];
    const int64_t tiles = (ncol + 32 * W - 1) / (32 * W);
    const int64_t outer = last == 0 ? 1 : a.n_out / ncol;
    const int64_t total = tiles * outer;
    const int split = blockIdx.y;
    // 64-bit divisions in front of the first load cost a 10 us kernel several per cent: 32-bit when the operands allow it,
    // none at all for a 2-d input (one kept axis: outer == 1)
    const int64_t chunk = a.n < ((int64_t)1 << 31) ? (int64_t)(((uint32_t)a.n + (uint32_t)a.splits - 1u) / (uint32_t)a.splits)
                                                    : (a.n + a.splits - 1) / a.splits;
    const int64_t r0 = (int64_t)split * chunk;
    const int64_t r1 = r0 + chunk < a.n ? r0 + chunk : a.n;
    for (int64_t vb = blockIdx.x; vb < total; vb += gridDim.x) {
        const int64_t o = outer == 1 ? 0 : vb / tiles;
        const int64_t tile = vb - o * tiles;
        // offsets of this outer index (all kept axes but the last)
        int64_t in_off = 0, out_off = 0, out2_off = 0;
        {
            int64_t lin = o;
            for (int d = last - 1; d >= 0; --d) {
                const int64_t q = lin / a.k.dim[d];
                const int64_t i = lin - q * a.k.dim[d];
                lin = q;
                in_off += i * a.k.in_stride[d];
                out_off += i * a.k.out_stride[d];
                out2_off += i * a.k.out2_stride[d];
            }
        }
        const int64_t c0 = tile * (32 * W) + (int64_t)tx * W;
        Acc<T, R> acc[W];
#pragma unroll
        for (int e = 0; e < W; ++e) acc[e].reset();
        const T *src = reinterpret_cast<const T *>(a.in) + in_off + c0;
        if (c0 < ncol) {  // ncol % W == 0 when W > 1 (host check), so a vector is all-in or all-out
            int64_t r = r0 + ty;
            constexpr int U = (R == R_WELFORD) ? NDB_RED_COLS_WU : NDB_RED_COLS_U;  // rows in flight per thread
            if constexpr (W > 1) {
                if constexpr (PEER) {
                    // The tile geometry of the fused sharded reduction must be the same on every rank, so a rank whose
                    // shard happens to start off a 16-byte boundary keeps W columns per thread and loads them one by one.
                    if (a.peer_scalar) {
                        for (; r < r1; r += 8) {
#pragma unroll
                            for (int e = 0; e < W; ++e) acc[e].add(src[r * a.s + e]);
                        }
                    }
                }
                for (; r + 8 * (U - 1) < r1; r += 8 * U) {
                    V v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = *reinterpret_cast<const V *>(src + (r + 8 * u) * a.s);
#pragma unroll
                    for (int e = 0; e < W; ++e) {
                        T x[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) x[u] = v[u].v[e];
                        acc[e].template add_batch<U>(x);
                    }
                }
                for (; r < r1; r += 8) {
                    V v0 = *reinterpret_cast<const V *>(src + r * a.s);
#pragma unroll
                    for (int e = 0; e < W; ++e) acc[e].add(v0.v[e]);
                }
            } else {
                for (; r + 8 * (U - 1) < r1; r += 8 * U) {
                    T x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u] = src[(r + 8 * u) * a.s];
                    acc[0].template add_batch<U>(x);
                }
                for (; r < r1; r += 8) acc[0].add(src[r * a.s]);
            }
        }
#pragma unroll
        for (int e = 0; e < W; ++e) smem[ty][tx * W + e] = acc[e];
        __syncthreads();
        // 32*W columns, merged over ty in fixed order by the first 32*W threads
        const int t = ty * 32 + tx;
        Acc<T, R> r;
        r.reset();
        const int64_t c = tile * (32 * W) + t;
        const bool mine = t < 32 * W && c < ncol;
        if (t < 32 * W) {
            r = smem[0][t];
#pragma unroll
            for (int j = 1; j < 8; ++j) r.merge(smem[j][t]);
            if (c < ncol) {
                if (a.splits == 1) {
                    if constexpr (!PEER) store_result<T, R, FIN>(a, r, out_off + c * a.k.out_stride[last], out2_off + c * a.k.out2_stride[last]);
                } else {
                    reinterpret_cast<Acc<T, R> *>(a.partial)[(int64_t)split * a.n_out + o * ncol + c] = r;
                }
            }
        }
        bool finisher = PEER && a.splits == 1;  // does this block hold the tile's complete local state in `r`?
        if (a.splits > 1 && a.tickets != nullptr) {
            // One launch instead of two: the block that takes the last ticket of this column tile merges all splits'
            // partial states IN SPLIT ORDER (so the result does not depend on which block came last) and stores.
            __shared__ unsigned ticket;
            __threadfence();  // this block's partials are visible device-wide before its ticket is
            __syncthreads();
            if (t == 0) ticket = atomicAdd(a.tickets + vb, 1u);
            __syncthreads();
            if (ticket == (unsigned)a.splits - 1) {
                __threadfence();
                if (mine) {
                    const Acc<T, R> *p = reinterpret_cast<const Acc<T, R> *>(a.partial) + o * ncol + c;
                    r = merge_splits(p, a.splits, a.n_out);
                    if constexpr (!PEER) store_result<T, R, FIN>(a, r, out_off + c * a.k.out_stride[last], out2_off + c * a.k.out2_stride[last]);
                }
                if (t == 0) a.tickets[vb] = 0;  // ready for the next launch
                finisher = PEER;
            }
        }
        if constexpr (PEER) {
            // Sharded reduction, still one launch: the block that completed this tile locally stores its states into
            // slot[rank] of EVERY rank's window (NVLink), bumps every rank's counter for this tile, waits until all
            // ranks' states for the tile have landed here and merges them in rank order before finalizing.
            if (finisher) {  // block-uniform
                const int64_t col = o * ncol + c;
                if (mine)
                    for (int q = 0; q < a.peer.world; ++q) reinterpret_cast<Acc<T, R> *>(peer_slot(a.peer, q, a.peer.rank))[col] = r;
                __threadfence_system();
                __syncthreads();
                if (t < a.peer.world)
                    atomicAdd_system(reinterpret_cast<unsigned long long *>(a.peer.win[t]) + 1 + vb, 1ull);
                if (t == 0) {
                    const unsigned long long done = a.tile_epoch[vb];
                    peer_spin(a.peer, 1 + (int)vb, (done + 1) * (unsigned long long)a.peer.world);
                    a.tile_epoch[vb] = done + 1;
                }
                __syncthreads();
                if (mine) {
                    Acc<T, R> tot = ldcg_state(reinterpret_cast<const Acc<T, R> *>(peer_slot(a.peer, a.peer.rank, 0)) + col);
                    for (int q = 1; q < a.peer.world; ++q)
                        tot.merge(ldcg_state(reinterpret_cast<const Acc<T, R> *>(peer_slot(a.peer, a.peer.rank, q)) + col));
                    store_result<T, R, FIN>(a, tot, out_off + c * a.k.out_stride[last], out2_off + c * a.k.out2_stride[last]);
                }
            }
        }
        __syncthreads();
    }
}

// merge per-split partial states in split order, finalize, scatter to the output view
template <typename T, int R, int FIN>
__global__ void __launch_bounds__(256) red_finish_cols(RedArgs a)
{
    pdl_enter();
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= a.n_out) return;
    const Acc<T, R> *p = reinterpret_cast<const Acc<T, R> *>(a.partial);
    Acc<T, R> r = merge_splits(p + o, a.splits, a.n_out);
    int64_t in_off, out_off, out2_off;
    kept_offsets(a.k, o, in_off, out_off, out2_off);
    store_result<T, R, FIN>(a, r, out_off, out2_off);
}

// ---- generic: one thread per output ------------------------------------------------------------------
template <typename T, int R, int FIN>
__global__ void __launch_bounds__(256) red_generic(RedArgs a)
{
    pdl_enter();
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < a.n_out; o += (int64_t)gridDim.x * blockDim.x) {
        int64_t in_off, out_off, out2_off;
        kept_offsets(a.k, o, in_off, out_off, out2_off);
        const T *src = reinterpret_cast<const T *>(a.in) + in_off;
        Acc<T, R> acc;
        acc.reset();
        for (int64_t i = 0; i < a.n; ++i) acc.add(src[i * a.s]);
        store_result<T, R, FIN>(a, acc, out_off, out2_off);
    }
}

// ---- whole-view reduction ---------------------------------------------------------------------------
struct AllArgs {
    const char *in;
    void *partial;  // [gridDim.x] states
    int ndim;       // collapsed rank
    int64_t dim[NDB_MAX_NDIM];
    int64_t stride[NDB_MAX_NDIM];
    int64_t total;
    int64_t inner_chunks;  // chunks of 1024 elements per row
    int64_t n_rows;
};

// dense unit-stride run, 128-bit loads, 4 vectors in flight per thread
template <typename T, int R>
__global__ void __launch_bounds__(256) red_all_flat(AllArgs a)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    __shared__ Acc<T, R> smem[32];
    const T *src = reinterpret_cast<const T *>(a.in);
    // peel to 16-byte alignment
    const int64_t mis = (int64_t)(((uintptr_t)src % 16) / sizeof(T));
    const int64_t head = mis ? ((N - mis) < a.total ? (N - mis) : a.total) : 0;
    const V *vsrc = reinterpret_cast<const V *>(src + head);
    const int64_t nv = (a.total - head) / N;
    Acc<T, R> acc;
    acc.reset();
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < nv; i += 4 * stride) {
        V v0 = vsrc[i], v1 = vsrc[i + stride], v2 = vsrc[i + 2 * stride], v3 = vsrc[i + 3 * stride];
        T x[4 * N];
#pragma unroll
        for (int e = 0; e < N; ++e) {
            x[e] = v0.v[e];
            x[N + e] = v1.v[e];
            x[2 * N + e] = v2.v[e];
            x[3 * N + e] = v3.v[e];
        }
        acc.template add_batch<4 * N>(x);
    }
    for (; i < nv; i += stride) {
        V v0 = vsrc[i];
        acc.template add_batch<N>(v0.v);
    }
    if (blockIdx.x == 0) {  // ragged head and tail
        const int64_t tail0 = head + nv * N;
        if (threadIdx.x < head) acc.add(src[threadIdx.x]);
        const int64_t t = tail0 + threadIdx.x;
        if (t < a.total) acc.add(src[t]);
    }
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) reinterpret_cast<Acc<T, R> *>(a.partial)[blockIdx.x] = acc;
}

// collapsed N-d view: virtual blocks of (row, 1024-element chunk of the inner axis)
template <typename T, int R>
__global__ void __launch_bounds__(256) red_all_rows(AllArgs a)
{
    pdl_enter();
    __shared__ Acc<T, R> smem[32];
    const T *src = reinterpret_cast<const T *>(a.in);
    const int inner = a.ndim - 1;
    const int64_t len = a.dim[inner], st = a.stride[inner];
    const int64_t total_vb = a.n_rows * a.inner_chunks;
    Acc<T, R> acc;
    acc.reset();
    for (int64_t vb = blockIdx.x; vb < total_vb; vb += gridDim.x) {
        const int64_t row = vb / a.inner_chunks;
        const int64_t ch = vb - row * a.inner_chunks;
        int64_t off = 0, lin = row;
        for (int d = inner - 1; d >= 0; --d) {
            const int64_t q = lin / a.dim[d];
            off += (lin - q * a.dim[d]) * a.stride[d];
            lin = q;
        }
        const int64_t i0 = ch * 1024 + threadIdx.x;
        T x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + j * 256;
            if (i < len) x[j] = src[off + i * st];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + j * 256;
            if (i < len) acc.add(x[j]);
        }
    }
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) reinterpret_cast<Acc<T, R> *>(a.partial)[blockIdx.x] = acc;
}

// one block merges the per-block partials in index order and writes the scalar
template <typename T, int R>
__global__ void __launch_bounds__(256) red_all_finish(const void *partial, int n, T *out, T dof, int want_std)
{
    pdl_enter();
    __shared__ Acc<T, R> smem[32];
    const Acc<T, R> *p = reinterpret_cast<const Acc<T, R> *>(partial);
    Acc<T, R> acc;
    acc.reset();
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc.merge(p[i]);
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) {
        if constexpr (R == R_WELFORD) {
            T v = acc.s.m2 / dof;
            if (want_std) v = std::is_same<T, float>::value ? (T)sqrtf((float)v) : (T)sqrt((double)v);
            *out = v;
        } else {
            *out = acc.s;
        }
    }
}

// The sharded sum()/product()/min/max: the block that finishes this rank's tree also does the exchange (peer.cuh): its
// scalar goes into slot[rank] of every rank's window over NVLink, and once all ranks have arrived the slots are merged in
// rank order.  Local reduction + "all-reduce" of one element cost the same two launches as the local reduction alone.
template <typename T, int R>
__global__ void __launch_bounds__(256) red_all_finish_peer(const void *partial, int n, T *out, PeerArgs pa)
{
    pdl_wait_only();
    __shared__ Acc<T, R> smem[32];
    const Acc<T, R> *p = reinterpret_cast<const Acc<T, R> *>(partial);
    Acc<T, R> acc;
    acc.reset();
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc.merge(p[i]);
    acc = block_reduce(acc, smem);
    __shared__ T total;
    if (threadIdx.x == 0) total = acc.s;
    __syncthreads();
    if ((int)threadIdx.x < pa.world) *reinterpret_cast<T *>(peer_slot(pa, threadIdx.x, pa.rank)) = total;
    peer_arrive(pa);
    peer_wait(pa);
    if (threadIdx.x == 0) {
        // every slot holds a value (an empty shard sent the identity / the type's extreme), so min / max states are seeded
        Acc<T, R> tot;
        tot.reset();
        tot.s = __ldcg(reinterpret_cast<const T *>(peer_slot(pa, pa.rank, 0)));
        tot.valid = true;
        for (int r = 1; r < pa.world; ++r) {
            Acc<T, R> b;
            b.reset();
            b.s = __ldcg(reinterpret_cast<const T *>(peer_slot(pa, pa.rank, r)));
            b.valid = true;
            tot.merge(b);
        }
        *out = tot.s;
    }
}

// dot product partials: x[i*sx] * y[i*sy], separately rounded product then sum (numeric_util.rs:56)
template <typename T>
__global__ void __launch_bounds__(256) dot_partials(const T *x, int64_t sx, const T *y, int64_t sy, int64_t n, void *partial)
{
    pdl_enter();
    __shared__ Acc<T, R_SUM> smem[32];
    Acc<T, R_SUM> acc;
    acc.reset();
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const T a0 = x[i * sx], a1 = x[(i + stride) * sx], a2 = x[(i + 2 * stride) * sx], a3 = x[(i + 3 * stride) * sx];
        const T b0 = y[i * sy], b1 = y[(i + stride) * sy], b2 = y[(i + 2 * stride) * sy], b3 = y[(i + 3 * stride) * sy];
        acc.add(w_mul(a0, b0));
        acc.add(w_mul(a1, b1));
        acc.add(w_mul(a2, b2));
        acc.add(w_mul(a3, b3));
    }
    for (; i < n; i += stride) acc.add(w_mul(x[i * sx], y[i * sy]));
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) reinterpret_cast<Acc<T, R_SUM> *>(partial)[blockIdx.x] = acc;
}

// dense, 16-byte aligned operands: 128-bit loads, 4 vector pairs in flight per thread
template <typename T>
__global__ void __launch_bounds__(256) dot_partials_vec(const T *x, const T *y, int64_t n, void *partial)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    __shared__ Acc<T, R_SUM> smem[32];
    Acc<T, R_SUM> acc;
    acc.reset();
    const V *vx = reinterpret_cast<const V *>(x), *vy = reinterpret_cast<const V *>(y);
    const int64_t nv = n / N;
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    T part[4] = {T(0), T(0), T(0), T(0)};
    for (; i + 3 * stride < nv; i += 4 * stride) {
        V a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = vx[i + u * stride];
            b[u] = vy[i + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < N; ++e) part[u] = w_add(part[u], w_mul(a[u].v[e], b[u].v[e]));
    }
    for (; i < nv; i += stride) {
        const V a = vx[i], b = vy[i];
#pragma unroll
        for (int e = 0; e < N; ++e) part[0] = w_add(part[0], w_mul(a.v[e], b.v[e]));
    }
    acc.add(w_add(w_add(part[0], part[1]), w_add(part[2], part[3])));
    if (blockIdx.x == 0) {  // ragged tail
        const int64_t t = nv * N + threadIdx.x;
        if (t < n) acc.add(w_mul(x[t], y[t]));
    }
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) reinterpret_cast<Acc<T, R_SUM> *>(partial)[blockIdx.x] = acc;
}

// Fused Zip + reduction over two dense, 16-byte aligned producers: reduce_R(op(x[i], y[i])).  This is the flat case
// of Zip::all / Zip::any (predicate op, MIN / MAX) and of an additive Zip::fold (arithmetic op, SUM): no scratch array.
template <typename T, int OP, int R>
__global__ void __launch_bounds__(256) zip2_reduce_flat(const T *x, const T *y, int64_t n, void *partial)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    __shared__ Acc<T, R> smem[32];
    Acc<T, R> acc;
    acc.reset();
    const V *vx = reinterpret_cast<const V *>(x), *vy = reinterpret_cast<const V *>(y);
    const int64_t nv = n / N;
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < nv; i += 4 * stride) {
        V a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = vx[i + u * stride];
            b[u] = vy[i + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < N; ++e) {
                const T in[2] = {a[u].v[e], b[u].v[e]};
                acc.add(apply_op<T, OP, 2>(in, T(0), T(0), nullptr));
            }
    }
    for (; i < nv; i += stride) {
        const V a = vx[i], b = vy[i];
#pragma unroll
        for (int e = 0; e < N; ++e) {
            const T in[2] = {a.v[e], b.v[e]};
            acc.add(apply_op<T, OP, 2>(in, T(0), T(0), nullptr));
        }
    }
    if (blockIdx.x == 0) {  // ragged tail
        const int64_t t = nv * N + threadIdx.x;
        if (t < n) {
            const T in[2] = {x[t], y[t]};
            acc.add(apply_op<T, OP, 2>(in, T(0), T(0), nullptr));
        }
    }
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) reinterpret_cast<Acc<T, R> *>(partial)[blockIdx.x] = acc;
}

template <typename T, int OP, int R>
static int zip2_reduce_launch(ndb_ctx *ctx, const void *x, const void *y, int64_t n, void *out_host)
{
    const int max_blocks = ctx->sm_count * 8;
    NDB_TRY(ensure_workspace(ctx, (size_t)max_blocks * sizeof(Acc<T, R>) + 4096));
    void *dev_scalar = (char *)ctx->workspace + ctx->workspace_bytes - 64;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, (n / Vec128<T>::N + 1023) / 1024));
    launch_pdl(zip2_reduce_flat<T, OP, R>, blocks, 256, 0, ctx->stream, (const T *)x, (const T *)y, n, ctx->workspace);
    NDB_CHECK_LAUNCH();
    launch_pdl(red_all_finish<T, R>, 1, 256, 0, ctx->stream, ctx->workspace, blocks, (T *)dev_scalar, T(1), 0);
    NDB_CHECK_LAUNCH();
    NDB_CUDA(cudaMemcpyAsync(ctx->host_scalar, dev_scalar, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out_host, ctx->host_scalar, sizeof(T));
    return NDB_OK;
}

template <typename T>
static int zip2_reduce_typed(ndb_ctx *ctx, int op, int redop, const void *x, const void *y, int64_t n, void *out_host)
{
#define NDB_Z2R(OPC, RC, RT) \
    if (op == OPC && redop == RC) return zip2_reduce_launch<T, OPC, RT>(ctx, x, y, n, out_host);
    NDB_Z2R(NDB_OP_MUL, NDB_RED_SUM, R_SUM)
    NDB_Z2R(NDB_OP_ADD, NDB_RED_SUM, R_SUM)
    NDB_Z2R(NDB_OP_SUB, NDB_RED_SUM, R_SUM)
    NDB_Z2R(NDB_OP_EQ, NDB_RED_MIN, R_MIN) NDB_Z2R(NDB_OP_EQ, NDB_RED_MAX, R_MAX)
    NDB_Z2R(NDB_OP_NE, NDB_RED_MIN, R_MIN) NDB_Z2R(NDB_OP_NE, NDB_RED_MAX, R_MAX)
    NDB_Z2R(NDB_OP_LT, NDB_RED_MIN, R_MIN) NDB_Z2R(NDB_OP_LT, NDB_RED_MAX, R_MAX)
    NDB_Z2R(NDB_OP_LE, NDB_RED_MIN, R_MIN) NDB_Z2R(NDB_OP_LE, NDB_RED_MAX, R_MAX)
    NDB_Z2R(NDB_OP_GT, NDB_RED_MIN, R_MIN) NDB_Z2R(NDB_OP_GT, NDB_RED_MAX, R_MAX)
    NDB_Z2R(NDB_OP_GE, NDB_RED_MIN, R_MIN) NDB_Z2R(NDB_OP_GE, NDB_RED_MAX, R_MAX)
#undef NDB_Z2R
    return NDB_ERR_UNSUPPORTED;  // the caller falls back to scratch + reduction
}

// returns NDB_ERR_UNSUPPORTED (without setting an error) when no fused kernel exists for (op, redop)
int zip2_reduce_dense(ndb_ctx *ctx, int op, int redop, int dtype, const void *x, const void *y, int64_t n, void *out_host)
{
    switch (dtype) {
    case NDB_F32: return zip2_reduce_typed<float>(ctx, op, redop, x, y, n, out_host);
    case NDB_F64: return zip2_reduce_typed<double>(ctx, op, redop, x, y, n, out_host);
    case NDB_I32: return zip2_reduce_typed<int32_t>(ctx, op, redop, x, y, n, out_host);
    case NDB_I64: return zip2_reduce_typed<int64_t>(ctx, op, redop, x, y, n, out_host);
    }
    return NDB_ERR_UNSUPPORTED;
}

// Fused Zip + fold of the WRITTEN values over dense, 16-byte aligned producers: out[i] = op(in...[i]) is stored and, in the
// same pass, reduced.  This is `Zip::fold` whose closure also writes through its first producer -
//     Zip::from(&mut a).and(&b).and(&c).fold(0., |acc, a, &b, &c| { *a = b * c + *a; acc + *a })
// - i.e. BASELINE config 5 (elementwise, then the global sum) in ONE pass: 16 bytes per element instead of 20.
// Persistent grid (one resident wave), two vectors per producer in flight per thread.
template <typename T, int OP, int NIN, int R>
__global__ void __launch_bounds__(256) zip_fold_flat(FlatArgs<NIN + 1> a, T s0, T s1, int *flag, void *partial)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    constexpr int U = NDB_ZIP_FOLD_U;
    __shared__ Acc<T, R> smem[32];
    Acc<T, R> acc;
    acc.reset();
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + (U - 1) * stride < a.n_vec; i += U * stride) {
        V in[U][NIN];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int p = 0; p < NIN; ++p) in[u][p] = reinterpret_cast<const V *>(a.ptr[p + 1])[i + u * stride];
        T outv[U * N];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            V o;
#pragma unroll
            for (int e = 0; e < N; ++e) {
                T x[NIN];
#pragma unroll
                for (int p = 0; p < NIN; ++p) x[p] = in[u][p].v[e];
                o.v[e] = apply_op<T, OP, NIN>(x, s0, s1, flag);
                outv[u * N + e] = o.v[e];
            }
            reinterpret_cast<V *>(a.ptr[0])[i + u * stride] = o;
        }
        acc.template add_batch<U * N>(outv);
    }
    for (; i < a.n_vec; i += stride) {
        V o;
#pragma unroll
        for (int e = 0; e < N; ++e) {
            T x[NIN];
#pragma unroll
            for (int p = 0; p < NIN; ++p) x[p] = reinterpret_cast<const V *>(a.ptr[p + 1])[i].v[e];
            o.v[e] = apply_op<T, OP, NIN>(x, s0, s1, flag);
        }
        reinterpret_cast<V *>(a.ptr[0])[i] = o;
        acc.template add_batch<N>(o.v);
    }
    if (blockIdx.x == 0) {  // ragged tail (< N elements), one thread each
        const int64_t t = a.n_vec * N + threadIdx.x;
        if (t < a.n) {
            T x[NIN];
#pragma unroll
            for (int p = 0; p < NIN; ++p) x[p] = reinterpret_cast<const T *>(a.ptr[p + 1])[t];
            const T o = apply_op<T, OP, NIN>(x, s0, s1, flag);
            reinterpret_cast<T *>(a.ptr[0])[t] = o;
            acc.add(o);
        }
    }
    acc = block_reduce(acc, smem);
    if (threadIdx.x == 0) reinterpret_cast<Acc<T, R> *>(partial)[blockIdx.x] = acc;
}

template <typename T, int OP, int NIN, int R>
static int zip_fold_launch(ndb_ctx *ctx, const ZipPlan &z, const void *scalars, T *out_dev, bool peer)
{
    FlatArgs<NIN + 1> a;
    for (int p = 0; p <= NIN; ++p) a.ptr[p] = z.ptr[p];
    a.n = z.total;
    a.n_vec = z.total / Vec128<T>::N;
    T s0 = T(0), s1 = T(0);
    const int ns = op_scalars(OP);
    if (ns >= 1 && scalars) s0 = ((const T *)scalars)[0];
    if (ns >= 2 && scalars) s1 = ((const T *)scalars)[1];
    const int max_blocks = ctx->sm_count * (getenv("NDB_ZIP_FOLD_PER_SM") ? std::max(1, atoi(getenv("NDB_ZIP_FOLD_PER_SM"))) : 4);  // 2^32 f32 MUL_ADD + sum, per SM 4 / 6 / 8: 10.18 / 10.49 / 10.58 ms
    NDB_TRY(ensure_workspace(ctx, (size_t)max_blocks * sizeof(Acc<T, R>)));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, (a.n_vec + 511) / 512));
    launch_pdl(zip_fold_flat<T, OP, NIN, R>, blocks, 256, 0, ctx->stream, a, s0, s1, ctx->dev_flag, ctx->workspace);
    NDB_CHECK_LAUNCH();
    if (peer) {
        PeerArgs pa;
        NDB_TRY(peer_begin(ctx, 1, sizeof(T), &pa));
        launch_pdl(red_all_finish_peer<T, R>, 1, 256, 0, ctx->stream, ctx->workspace, blocks, out_dev, pa);
    } else {
        launch_pdl(red_all_finish<T, R>, 1, 256, 0, ctx->stream, ctx->workspace, blocks, out_dev, T(1), 0);
    }
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

template <typename T, int R>
static int zip_fold_op(ndb_ctx *ctx, const ZipPlan &z, const void *scalars, T *out_dev, bool peer)
{
    switch (z.op) {
    case NDB_OP_COPY: return zip_fold_launch<T, NDB_OP_COPY, 1, R>(ctx, z, scalars, out_dev, peer);
    case NDB_OP_ADD: return zip_fold_launch<T, NDB_OP_ADD, 2, R>(ctx, z, scalars, out_dev, peer);
    case NDB_OP_SUB: return zip_fold_launch<T, NDB_OP_SUB, 2, R>(ctx, z, scalars, out_dev, peer);
    case NDB_OP_MUL: return zip_fold_launch<T, NDB_OP_MUL, 2, R>(ctx, z, scalars, out_dev, peer);
    case NDB_OP_AXPY: return zip_fold_launch<T, NDB_OP_AXPY, 2, R>(ctx, z, scalars, out_dev, peer);
    case NDB_OP_MUL_ADD: return zip_fold_launch<T, NDB_OP_MUL_ADD, 3, R>(ctx, z, scalars, out_dev, peer);
    default: return NDB_ERR_UNSUPPORTED;
    }
}

template <typename T>
static int zip_fold_typed(ndb_ctx *ctx, const ZipPlan &z, int redop, const void *scalars, void *out_dev, bool peer)
{
    switch (redop) {
    case NDB_RED_SUM: return zip_fold_op<T, R_SUM>(ctx, z, scalars, (T *)out_dev, peer);
    case NDB_RED_MIN: return zip_fold_op<T, R_MIN>(ctx, z, scalars, (T *)out_dev, peer);
    case NDB_RED_MAX: return zip_fold_op<T, R_MAX>(ctx, z, scalars, (T *)out_dev, peer);
    default: return NDB_ERR_UNSUPPORTED;
    }
}

// NDB_ERR_UNSUPPORTED (no error text) when no fused kernel covers (plan, op, redop): the caller runs Zip, then the reduction
int zip_fold_dense(ndb_ctx *ctx, const ZipPlan &z, int redop, const void *scalars, void *out_dev, bool peer)
{
    if (z.kernel != NDB_KERNEL_FLAT_VEC || z.n_parts != op_arity(z.op) + 1) return NDB_ERR_UNSUPPORTED;
    switch (z.dtype) {
    case NDB_F32: return zip_fold_typed<float>(ctx, z, redop, scalars, out_dev, peer);
    case NDB_F64: return zip_fold_typed<double>(ctx, z, redop, scalars, out_dev, peer);
    case NDB_I32: return zip_fold_typed<int32_t>(ctx, z, redop, scalars, out_dev, peer);
    case NDB_I64: return zip_fold_typed<int64_t>(ctx, z, redop, scalars, out_dev, peer);
    }
    return NDB_ERR_UNSUPPORTED;
}

// ======================================================================================================
// host side

// Collapse a read-only view for a whole-view reduction: drop unit axes, make strides positive (order is
// free for a commutative reduction up to the stated tolerance; exact for integers), sort, merge.
static void collapse_for_all(const ndb_view *v, AllArgs *a)
{
    int n = 0;
    int64_t d[NDB_MAX_NDIM], s[NDB_MAX_NDIM];
    int64_t off = 0;
    for (int i = 0; i < v->ndim; ++i) {
        if (v->dim[i] == 1) continue;
        d[n] = v->dim[i];
        s[n] = v->stride[i];
        if (s[n] < 0) {
            off += (d[n] - 1) * s[n];
            s[n] = -s[n];
        }
        ++n;
    }
    int order[NDB_MAX_NDIM];
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order, order + n, [&](int x, int y) { return s[x] > s[y]; });
    int m = 0;
    int64_t rd[NDB_MAX_NDIM], rs[NDB_MAX_NDIM];
    for (int k = n - 1; k >= 0; --k) {  // innermost outwards
        const int i = order[k];
        if (m > 0 && s[i] == rd[m - 1] * rs[m - 1]) rd[m - 1] *= d[i];
        else {
            rd[m] = d[i];
            rs[m] = s[i];
            ++m;
        }
    }
    if (m == 0) {
        rd[0] = 1;
        rs[0] = 1;
        m = 1;
    }
    a->ndim = m;
    a->total = 1;
    for (int k = 0; k < m; ++k) {
        a->dim[k] = rd[m - 1 - k];
        a->stride[k] = rs[m - 1 - k];
        a->total *= a->dim[k];
    }
    a->in = (const char *)v->ptr + off * (int64_t)dtype_size(v->dtype);
}

template <typename T, int R>
static int reduce_all_typed(ndb_ctx *ctx, const ndb_view *v, T *out_dev, T dof = T(1), int want_std = 0, bool peer = false)
{
    AllArgs a;
    memset(&a, 0, sizeof(a));
    collapse_for_all(v, &a);
    // one resident wave: the pairwise-batch kernel needs 32 registers, so eight 256-thread blocks fit per SM; measured
    // 2/3/4/6/8 per SM: 3.2/4.5/5.5/6.6/6.9 TB/s (C3 sum, f64 16384^2), 4.3/5.7/6.4/6.8/6.9 TB/s (f32)
    const int max_blocks = ctx->sm_count * (getenv("NDB_RED_ALL_PER_SM") ? std::max(1, atoi(getenv("NDB_RED_ALL_PER_SM"))) : 8);
    NDB_TRY(ensure_workspace(ctx, (size_t)max_blocks * sizeof(Acc<T, R>)));
    a.partial = ctx->workspace;
    int blocks;
    if (a.ndim == 1 && a.stride[0] == 1) {
        const int64_t nv = a.total / Vec128<T>::N;
        blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, (nv + 1023) / 1024));
        launch_pdl(red_all_flat<T, R>, blocks, 256, 0, ctx->stream, a);
    } else {
        const int inner = a.ndim - 1;
        a.inner_chunks = (a.dim[inner] + 1023) / 1024;
        a.n_rows = a.total / a.dim[inner];
        blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, a.n_rows * a.inner_chunks));
        launch_pdl(red_all_rows<T, R>, blocks, 256, 0, ctx->stream, a);
    }
    NDB_CHECK_LAUNCH();
    if constexpr (R != R_WELFORD) {
        if (peer) {
            PeerArgs pa;
            NDB_TRY(peer_begin(ctx, 1, sizeof(T), &pa));
            launch_pdl(red_all_finish_peer<T, R>, 1, 256, 0, ctx->stream, a.partial, blocks, out_dev, pa);
            NDB_CHECK_LAUNCH();
            return NDB_OK;
        }
    }
    launch_pdl(red_all_finish<T, R>, 1, 256, 0, ctx->stream, a.partial, blocks, out_dev, dof, want_std);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

template <typename T>
static int reduce_all_dispatch(ndb_ctx *ctx, int redop, const ndb_view *v, void *out_dev, bool peer = false)
{
    switch (redop) {
    case NDB_RED_SUM: return reduce_all_typed<T, R_SUM>(ctx, v, (T *)out_dev, T(1), 0, peer);
    case NDB_RED_PROD: return reduce_all_typed<T, R_PROD>(ctx, v, (T *)out_dev, T(1), 0, peer);
    case NDB_RED_MIN: return reduce_all_typed<T, R_MIN>(ctx, v, (T *)out_dev, T(1), 0, peer);
    case NDB_RED_MAX: return reduce_all_typed<T, R_MAX>(ctx, v, (T *)out_dev, T(1), 0, peer);
    default: break;
    }
    set_error("reduce_all: redop %d is not a whole-array reduction", redop);
    return NDB_ERR_UNSUPPORTED;
}

static int64_t view_len(const ndb_view *v)
{
    int64_t n = 1;
    for (int i = 0; i < v->ndim; ++i) n *= v->dim[i];
    return n;
}

// identity written for empty inputs: sum -> 0, product -> 1 (tests/numeric.rs:57-75).  min / max have none in ndarray
// (an empty fold returns its init); across ranks an EMPTY SHARD of a non-empty array still has to take part in the
// exchange, and contributes the element type's extreme value (+inf / MAX for min, -inf / MIN for max).
int write_identity(ndb_ctx *ctx, int dtype, int redop, void *out_dev)
{
    union { float f; double d; int32_t i; int64_t l; } u;
    memset(&u, 0, sizeof(u));
    switch (redop) {
    case NDB_RED_PROD:
        switch (dtype) {
        case NDB_F32: u.f = 1.f; break;
        case NDB_F64: u.d = 1.0; break;
        case NDB_I32: u.i = 1; break;
        default: u.l = 1; break;
        }
        break;
    case NDB_RED_MIN:
        switch (dtype) {
        case NDB_F32: u.f = INFINITY; break;
        case NDB_F64: u.d = INFINITY; break;
        case NDB_I32: u.i = INT32_MAX; break;
        default: u.l = INT64_MAX; break;
        }
        break;
    case NDB_RED_MAX:
        switch (dtype) {
        case NDB_F32: u.f = -INFINITY; break;
        case NDB_F64: u.d = -INFINITY; break;
        case NDB_I32: u.i = INT32_MIN; break;
        default: u.l = INT64_MIN; break;
        }
        break;
    default: break;  // sum: all-zero bits
    }
    // the value travels as a kernel parameter (a staged copy from ctx->host_scalar could be overtaken by the next call)
    ndb_view one;
    memset(&one, 0, sizeof(one));
    one.ptr = out_dev;
    one.dtype = dtype;
    one.ndim = 1;
    one.dim[0] = 1;
    one.stride[0] = 1;
    return ndb_zip(ctx, NDB_OP_FILL, &u, 1, &one);
}

static int reduce_all_dev_impl(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_dev, bool peer = false)
{
    NDB_TRY(validate_view(in, "reduce_all"));
    if (view_len(in) == 0) {
        if (peer) {
            // an empty shard still has to take part in the exchange (returning here would leave the other ranks waiting):
            // it contributes the identity, for min / max the type's extreme
            NDB_TRY(write_identity(ctx, in->dtype, redop, out_dev));
            return ndb_peer_allreduce(ctx, redop, in->dtype, out_dev, out_dev, 1);
        }
        if (redop == NDB_RED_SUM || redop == NDB_RED_PROD) return write_identity(ctx, in->dtype, redop, out_dev);
        set_error("reduce_all: min/max of an empty array");
        return NDB_ERR_EMPTY;
    }
    switch (in->dtype) {
    case NDB_F32: return reduce_all_dispatch<float>(ctx, redop, in, out_dev, peer);
    case NDB_F64: return reduce_all_dispatch<double>(ctx, redop, in, out_dev, peer);
    case NDB_I32: return reduce_all_dispatch<int32_t>(ctx, redop, in, out_dev, peer);
    case NDB_I64: return reduce_all_dispatch<int64_t>(ctx, redop, in, out_dev, peer);
    }
    return NDB_ERR_UNSUPPORTED;
}

// ---- axis reductions ------------------------------------------------------------------------------
template <typename T, int R, int FIN>
static int reduce_axis_launch(ndb_ctx *ctx, RedArgs &a, const ZipPlan &z)
{
    cudaStream_t st = ctx->stream;
    const int64_t esz = sizeof(T);
    const int64_t vecn = Vec128<T>::N;
    const int last = a.k.n - 1;
    int64_t min_kept = INT64_MAX;  // smallest non-zero |stride| among kept axes of the input
    for (int d = 0; d < a.k.n; ++d)
        if (a.k.dim[d] > 1) min_kept = std::min<int64_t>(min_kept, llabs(a.k.in_stride[d]));
    const int64_t as = llabs(a.s);
    // peer mode: the choice must not depend on this rank's row count (a rank holding one row, or none, still has to take
    // part in the exchange with the same tile geometry as the others)
    const bool lanes = a.n_out == 1 || (!a.peer_on && a.n == 1) || as <= min_kept;
    if (a.peer_on && (lanes || !(a.k.in_stride[last] == 1 && a.k.dim[last] >= 32))) return NDB_ERR_UNSUPPORTED;
    if (lanes) {
        bool vec = a.s == 1 && a.n % vecn == 0 && ((uintptr_t)a.in % 16) == 0 && a.n >= 4 * vecn;
        for (int d = 0; d < a.k.n && vec; ++d) vec = a.k.dim[d] == 1 || a.k.in_stride[d] % vecn == 0;
        const bool warp = a.n <= (getenv("NDB_RED_WARP_MAX") ? atoll(getenv("NDB_RED_WARP_MAX")) : 8192);
        // warp-per-lane: smaller blocks when 8 lanes per block would leave the SMs unevenly loaded (4096 lanes = 3.46
        // blocks of 8 per SM: the SMs holding 4 set the pace)
        int wpb = 8;
        // warps per lane (NDB_RED_LANES_P): two or four warps can share a lane and meet in shared memory.  Measured on
        // C1 (4096 lanes of 16 KiB, cold, back to back): P = 1 / 2 / 4 -> 13.3 / 14.2 / 14.5 us, so one warp per lane stays
        // the default; the knob is kept for shapes with few, long lanes.
        int P = 1;
        if (warp && getenv("NDB_RED_LANES_P")) P = std::max(1, std::min(8, atoi(getenv("NDB_RED_LANES_P"))));
        a.lane_warps = P;
        if (getenv("NDB_RED_WPB")) wpb = atoi(getenv("NDB_RED_WPB"));
        else while (wpb > 2 && wpb > P && (a.n_out * P + wpb - 1) / wpb < (int64_t)ctx->sm_count * 6) wpb >>= 1;
        if (wpb < P || wpb % P) wpb = P;
        const int64_t groups = warp ? (a.n_out * P + wpb - 1) / wpb : a.n_out;
        unsigned blocks = (unsigned)std::min<int64_t>(groups, (int64_t)1 << 30);
        if (getenv("NDB_RED_LANES_BLOCKS")) blocks = std::min<unsigned>(blocks, (unsigned)atoi(getenv("NDB_RED_LANES_BLOCKS")));
        if (warp) {
            if (vec) launch_pdl(red_lanes<T, R, FIN, true, true>, blocks, 32 * wpb, 0, st, a);
            else launch_pdl(red_lanes<T, R, FIN, true, false>, blocks, 32 * wpb, 0, st, a);
        } else {
            if (vec) launch_pdl(red_lanes<T, R, FIN, false, true>, blocks, 256, 0, st, a);
            else launch_pdl(red_lanes<T, R, FIN, false, false>, blocks, 256, 0, st, a);
        }
        NDB_CHECK_LAUNCH();
        return NDB_OK;
    }
    if (a.k.in_stride[last] == 1 && a.k.dim[last] >= 32) {
        const int64_t ncol = a.k.dim[last];
        const bool aligned = ((uintptr_t)a.in % 16) == 0;
        bool vec = ncol % vecn == 0 && (aligned || a.peer_on) && a.s % vecn == 0;  // peer: alignment is rank-local, see peer_scalar
        for (int d = 0; d < last && vec; ++d) vec = a.k.in_stride[d] % vecn == 0;
        a.peer_scalar = a.peer_on && !aligned;
        const int w = vec ? (int)vecn : 1;
        const int64_t tiles = (ncol + 32 * w - 1) / (32 * w);
        const int64_t total = tiles * (a.n_out / ncol);
        // Large problems: enough blocks to fill the machine several deep (measured best 8 per SM for the plain reducers, 6
        // for Welford); every split keeps >= 64 rows.  A problem of at most 128 MiB (C1: 4096^2 f32) is fastest as ONE
        // resident wave of at most two fat blocks per SM (ncu durations for C1, splits 5/8/9/10/14/19/37:
        // 18.6 / 15.6 / 15.8 / 16.2 / 20.2 / 19.7 / 21.6 us), so there the split count is rounded DOWN to 2 per SM.
        const bool small = (double)a.n * (double)a.n_out * sizeof(T) <= 128.0 * 1024 * 1024;
        const int per_sm = getenv("NDB_RED_COLS_PER_SM") ? atoi(getenv("NDB_RED_COLS_PER_SM")) : (small ? 2 : (R == R_WELFORD ? 6 : 8));
        int64_t want = small ? std::max<int64_t>(1, (int64_t)ctx->sm_count * per_sm / total) : ((int64_t)ctx->sm_count * per_sm + total - 1) / total;
        int64_t max_splits = std::max<int64_t>(1, a.n / 64);
        int splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, max_splits), 256));
        if (getenv("NDB_RED_COLS_SPLITS")) splits = (int)std::max<int64_t>(1, std::min<int64_t>(atoi(getenv("NDB_RED_COLS_SPLITS")), max_splits));
        a.splits = splits;
        a.tickets = nullptr;
        if (splits > 1) {
            NDB_TRY(ensure_workspace(ctx, (size_t)splits * (size_t)a.n_out * sizeof(Acc<T, R>)));
            a.partial = ctx->workspace;
            if (total <= NDB_TICKETS && !getenv("NDB_RED_TWO_KERNELS")) a.tickets = ctx->tickets;
        }
        dim3 grid((unsigned)std::min<int64_t>(total, (int64_t)1 << 30), (unsigned)splits, 1), block(32, 8, 1);
        if (a.peer_on) {
            // Fused exchange (see red_cols<PEER>): every tile's finishing block waits for the other ranks' finishers, so
            // the whole grid is launched COOPERATIVELY - the driver either makes every block co-resident or refuses the
            // launch - and is sized from the kernel's real occupancy.  A waiting block can therefore never keep a block it
            // (transitively) waits for off the machine, and the watchdog in peer_spin only fires if a peer process died.
            // Everything that decides eligibility (tile count, occupancy x SM count, slot size) is the same on every
            // rank for identically shaped shards; the split count is rank-local and free to differ.
            if constexpr (FIN == FIN_WELFORD_RAW) {
                return NDB_ERR_UNSUPPORTED;
            } else {
                auto kern = red_cols<T, R, FIN, Vec128<T>::N, true>;
                static std::atomic<int> occ_cache{0};
                int occ = occ_cache.load(std::memory_order_relaxed);
                if (occ == 0) {
                    NDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
                    occ_cache.store(occ, std::memory_order_relaxed);
                }
                const int64_t capacity = (int64_t)occ * ctx->sm_count;
                if (!vec || total > std::min<int64_t>(PEER_MAX_TILES, capacity) || total > NDB_TICKETS ||
                    (size_t)a.n_out * sizeof(Acc<T, R>) > ctx->peer_slot_bytes || !ctx->peer_tile_epoch)
                    return NDB_ERR_UNSUPPORTED;
                if (a.peer_on == 2) return NDB_OK;  // eligibility query only (ndb_reduce_axis_peer_ok): nothing is launched
                splits = (int)std::max<int64_t>(1, std::min<int64_t>(splits, capacity / total));
                a.splits = splits;
                a.tickets = splits > 1 ? ctx->tickets : nullptr;
                grid.y = (unsigned)splits;
                NDB_TRY(peer_begin(ctx, 0, (size_t)a.n_out * sizeof(Acc<T, R>), &a.peer));
                a.tile_epoch = ctx->peer_tile_epoch;
                void *kargs[1] = {(void *)&a};
                NDB_CUDA(cudaLaunchCooperativeKernel((const void *)kern, grid, block, kargs, 0, st));
                NDB_CHECK_LAUNCH();
                return NDB_OK;
            }
        }
        if (vec) launch_pdl(red_cols<T, R, FIN, Vec128<T>::N>, grid, block, 0, st, a);
        else launch_pdl(red_cols<T, R, FIN, 1>, grid, block, 0, st, a);
        NDB_CHECK_LAUNCH();
        if (splits > 1 && a.tickets == nullptr) {
            launch_pdl(red_finish_cols<T, R, FIN>, (unsigned)((a.n_out + 255) / 256), 256, 0, st, a);
            NDB_CHECK_LAUNCH();
        }
        return NDB_OK;
    }
    (void)esz;
    (void)z;
    const unsigned blocks = (unsigned)std::min<int64_t>((a.n_out + 255) / 256, (int64_t)1 << 30);
    launch_pdl(red_generic<T, R, FIN>, blocks, 256, 0, st, a);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

template <typename T>
static int reduce_axis_dispatch(ndb_ctx *ctx, int redop, RedArgs &a, const ZipPlan &z)
{
    switch (redop) {
    case NDB_RED_SUM: return reduce_axis_launch<T, R_SUM, FIN_NONE>(ctx, a, z);
    case NDB_RED_PROD: return reduce_axis_launch<T, R_PROD, FIN_NONE>(ctx, a, z);
    case NDB_RED_MIN: return reduce_axis_launch<T, R_MIN, FIN_NONE>(ctx, a, z);
    case NDB_RED_MAX: return reduce_axis_launch<T, R_MAX, FIN_NONE>(ctx, a, z);
    case NDB_RED_MEAN: return reduce_axis_launch<T, R_SUM, FIN_MEAN>(ctx, a, z);
    default: break;
    }
    if constexpr (std::is_floating_point<T>::value) {
        switch (redop) {
        case NDB_RED_VAR: return reduce_axis_launch<T, R_WELFORD, FIN_VAR>(ctx, a, z);
        case NDB_RED_STD: return reduce_axis_launch<T, R_WELFORD, FIN_STD>(ctx, a, z);
        case -1: return reduce_axis_launch<T, R_WELFORD, FIN_WELFORD_RAW>(ctx, a, z);
        default: break;
        }
    }
    set_error("reduce_axis: redop %d unsupported for this element type", redop);
    return NDB_ERR_UNSUPPORTED;
}

// Fill `out` with a constant through the Zip engine (empty-axis conventions).
static int fill_view(ndb_ctx *ctx, const ndb_view *out, double value)
{
    union { float f; double d; int32_t i; int64_t l; } u;
    switch (out->dtype) {
    case NDB_F32: u.f = (float)value; break;
    case NDB_F64: u.d = value; break;
    case NDB_I32: u.i = (int32_t)value; break;
    default: u.l = (int64_t)value; break;
    }
    return ndb_zip(ctx, NDB_OP_FILL, &u, 1, out);
}

static int reduce_axis_impl(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out,
                            const ndb_view *out2, double ddof, int peer = 0, int64_t n_global = 0)
{
    if (!ctx) {
        set_error("reduce_axis: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    NDB_TRY(validate_view(in, "reduce_axis input"));
    NDB_TRY(validate_view(out, "reduce_axis output"));
    if (axis < 0 || axis >= in->ndim) {
        set_error("axis %d is out of bounds for array of dimension %d", axis, in->ndim);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    if (out->dtype != in->dtype || (out2 && out2->dtype != in->dtype)) {
        set_error("reduce_axis: output dtype differs from input dtype");
        return NDB_ERR_UNSUPPORTED;
    }
    // `in` with the reduced axis removed
    ndb_view base = *in;
    for (int i = axis; i + 1 < in->ndim; ++i) {
        base.dim[i] = in->dim[i + 1];
        base.stride[i] = in->stride[i + 1];
    }
    base.ndim = in->ndim - 1;
    const int64_t n = in->dim[axis];
    if (peer && n_global < n) return NDB_ERR_UNSUPPORTED;
    if (peer && n_global == 0) return NDB_ERR_UNSUPPORTED;  // the empty-axis conventions are the unsharded call's business
    const int64_t n_stat = peer ? n_global : n;  // the length mean / var / std divide by: the whole axis, all shards
    const bool welford = redop == NDB_RED_VAR || redop == NDB_RED_STD || redop == -1;
    if (welford && !dtype_is_float(in->dtype)) {
        set_error("var_axis/std_axis need a float element type");
        return NDB_ERR_UNSUPPORTED;
    }
    double nT = (double)n_stat, dof = 0.0;
    if (redop == NDB_RED_VAR || redop == NDB_RED_STD) {
        // assert!(!(ddof < zero || ddof > n)) in the element type; NaN passes (impl_numeric.rs:408-413)
        if (in->dtype == NDB_F32) {
            const float d = (float)ddof, nf = (float)n_stat;
            if (d < 0.f || d > nf) {
                set_error("`ddof` must not be less than zero or greater than the length of the axis");
                return NDB_ERR_INVALID_ARG;
            }
            dof = (double)(nf - d);
            nT = (double)nf;
        } else {
            if (ddof < 0.0 || ddof > (double)n_stat) {
                set_error("`ddof` must not be less than zero or greater than the length of the axis");
                return NDB_ERR_INVALID_ARG;
            }
            dof = (double)n_stat - ddof;
        }
    }
    // the plan validates out/base shapes, drops unit axes, flips and merges jointly
    ndb_view parts[3] = {*out, base, out2 ? *out2 : *out};
    ZipPlan z;
    int st = make_zip_plan(out2 ? NDB_OP_ADD : NDB_OP_COPY, out2 ? 3 : 2, parts, &z);
    if (st == NDB_ERR_INCOMPATIBLE_SHAPE) set_error("reduce_axis: output shape must be the input shape without the axis");
    NDB_TRY(st);
    if (z.kernel == NDB_KERNEL_NOOP) return NDB_OK;  // no outputs at all
    DeviceGuard g(ctx->device);
    if (n == 0 && !peer) {  // (a rank whose SHARD is empty still takes part in the fused exchange, contributing no rows)
        // sum -> 0, product -> 1, mean -> None, var/std -> NaN (0/0), min/max -> nothing to fold
        if (redop == NDB_RED_SUM) return fill_view(ctx, out, 0.0);
        if (redop == NDB_RED_PROD) return fill_view(ctx, out, 1.0);
        if (redop == NDB_RED_VAR || redop == NDB_RED_STD) return fill_view(ctx, out, NAN);
        if (redop == -1) {
            NDB_TRY(fill_view(ctx, out, 0.0));
            return fill_view(ctx, out2, 0.0);
        }
        set_error("reduction of a zero-length axis has no value (ndarray returns None)");
        return NDB_ERR_EMPTY;
    }
    RedArgs a;
    memset(&a, 0, sizeof(a));
    a.out = z.ptr[0];
    a.in = z.ptr[1];
    a.out2 = out2 ? z.ptr[2] : nullptr;
    a.k.n = z.ndim;
    a.n_out = z.total;
    for (int d = 0; d < z.ndim; ++d) {
        a.k.dim[d] = z.dim[d];
        a.k.out_stride[d] = z.stride[0][d];
        a.k.in_stride[d] = z.stride[1][d];
        a.k.out2_stride[d] = out2 ? z.stride[2][d] : 0;
    }
    a.n = n;
    a.s = (n > 1 || peer) ? in->stride[axis] : 1;  // peer: kernel choice must not depend on this rank's row count
    a.nT = nT;
    a.dof = dof;
    a.splits = 1;
    a.peer_on = peer;
    // a backwards reduced axis is walked forwards from its low end (legal: commutative up to tolerance)
    if (a.s < 0) {
        a.in += (n - 1) * a.s * (int64_t)dtype_size(in->dtype);
        a.s = -a.s;
    }
    switch (in->dtype) {
    case NDB_F32: return reduce_axis_dispatch<float>(ctx, redop, a, z);
    case NDB_F64: return reduce_axis_dispatch<double>(ctx, redop, a, z);
    case NDB_I32: return reduce_axis_dispatch<int32_t>(ctx, redop, a, z);
    case NDB_I64: return reduce_axis_dispatch<int64_t>(ctx, redop, a, z);
    }
    return NDB_ERR_UNSUPPORTED;
}

// merge Welford shards: out = finalize(merge_s (n_s, mean_s, m2_s)) — Chan's formula in shard order
template <typename T>
__global__ void __launch_bounds__(256) welford_merge_kernel(int n_shards, const int64_t *counts, const T *const *means,
                                                            const T *const *m2s, int64_t n_out, T dof, int want_std, T *out)
{
    pdl_enter();
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    using Rd = Reducer<T, R_WELFORD>;
    typename Rd::State s = Rd::init();
    for (int k = 0; k < n_shards; ++k) {
        typename Rd::State t{(T)counts[k], means[k][o], m2s[k][o]};
        Rd::merge(s, t);
    }
    T v = s.m2 / dof;
    if (want_std) v = std::is_same<T, float>::value ? (T)sqrtf((float)v) : (T)sqrt((double)v);
    out[o] = v;
}

// shards gathered from all ranks (means / m2s are [world][n_out], counts [world], all on the device): merge in rank order
template <typename T>
__global__ void __launch_bounds__(256) welford_merge_gathered_kernel(int world, const int64_t *counts, const T *means, const T *m2s,
                                                                     int64_t n_out, int64_t pitch, T ddof, int want_std, T *out)
{
    pdl_enter();
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    using Rd = Reducer<T, R_WELFORD>;
    typename Rd::State s = Rd::init();
    for (int k = 0; k < world; ++k) {
        typename Rd::State t{(T)counts[k], means[(int64_t)k * pitch + o], m2s[(int64_t)k * pitch + o]};
        Rd::merge(s, t);
    }
    T v = s.m2 / (s.n - ddof);
    if (want_std) v = std::is_same<T, float>::value ? (T)sqrtf((float)v) : (T)sqrt((double)v);
    out[o] = v;
}

int welford_merge_gathered(ndb_ctx *ctx, int dtype, int world, const int64_t *counts_dev, const void *means_dev, const void *m2s_dev,
                           int64_t n_out, int64_t pitch, double ddof, int want_std, void *out_dev)
{
    if (n_out <= 0) return NDB_OK;
    const unsigned blocks = (unsigned)((n_out + 255) / 256);
    if (dtype == NDB_F32)
        launch_pdl(welford_merge_gathered_kernel<float>, blocks, 256, 0, ctx->stream, world, counts_dev, (const float *)means_dev,
                   (const float *)m2s_dev, n_out, pitch, (float)ddof, want_std, (float *)out_dev);
    else if (dtype == NDB_F64)
        launch_pdl(welford_merge_gathered_kernel<double>, blocks, 256, 0, ctx->stream, world, counts_dev, (const double *)means_dev,
                   (const double *)m2s_dev, n_out, pitch, ddof, want_std, (double *)out_dev);
    else return NDB_ERR_UNSUPPORTED;
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_reduce_all_dev(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_dev)
{
    if (!ctx || !out_dev) {
        set_error("reduce_all: null context or output (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    clear_error();
    DeviceGuard g(ctx->device);
    return reduce_all_dev_impl(ctx, redop, in, out_dev);
}

int ndb_reduce_all_peer(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_dev)
{
    if (!ctx || !out_dev) {
        set_error("reduce_all_peer: null context or output (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    clear_error();
    DeviceGuard g(ctx->device);
    return reduce_all_dev_impl(ctx, redop, in, out_dev, true);
}

int ndb_zip_fold_write(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts, int redop, void *out_dev,
                       int across_ranks)
{
    clear_error();
    if (!ctx || !out_dev || !parts) {
        set_error("zip_fold_write: null context, producers or output (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    if (redop < NDB_RED_SUM || redop > NDB_RED_MAX) {
        set_error("zip_fold_write: the accumulator must be sum, product, min or max");
        return NDB_ERR_UNSUPPORTED;
    }
    const bool peer = across_ranks != 0;
    if (peer && !(ctx->peer_world > 0 && ctx->peer_attached)) {
        set_error("zip_fold_write: across_ranks needs peer windows attached to this context");
        return NDB_ERR_INVALID_ARG;
    }
    ZipPlan z;
    NDB_TRY(make_zip_plan(op, n_parts, parts, &z));
    DeviceGuard g(ctx->device);
    if (z.kernel != NDB_KERNEL_NOOP) {
        const int st = zip_fold_dense(ctx, z, redop, scalars, out_dev, peer);
        if (st != NDB_ERR_UNSUPPORTED) return st;
        clear_error();
    }
    // general producers: the Zip, then the reduction of what it wrote (two passes over the output)
    NDB_TRY(ndb_zip(ctx, op, scalars, n_parts, parts));
    return reduce_all_dev_impl(ctx, redop, &parts[0], out_dev, peer);
}

int ndb_reduce_all(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_host)
{
    if (!ctx || !out_host) {
        set_error("reduce_all: null context or output (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    clear_error();
    DeviceGuard g(ctx->device);
    NDB_TRY(ensure_workspace(ctx, 4096));
    // the scalar lands at the END of the partials buffer's first page (partials use the front)
    void *dev_scalar = (char *)ctx->workspace + ctx->workspace_bytes - 64;
    NDB_TRY(reduce_all_dev_impl(ctx, redop, in, dev_scalar));
    NDB_CUDA(cudaMemcpyAsync(ctx->host_scalar, dev_scalar, dtype_size(in->dtype), cudaMemcpyDeviceToHost, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out_host, ctx->host_scalar, dtype_size(in->dtype));
    return NDB_OK;
}

// ArrayRef::var / std (src/numeric/impl_numeric.rs:180-249): Welford over every element, M2 / (n - ddof).
int ndb_var(ndb_ctx *ctx, const ndb_view *in, double ddof, int want_std, void *out_host)
{
    clear_error();
    if (!ctx || !out_host) {
        set_error("var: null context or output (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    NDB_TRY(validate_view(in, "var"));
    if (!dtype_is_float(in->dtype)) {
        set_error("var/std need a float element type");
        return NDB_ERR_UNSUPPORTED;
    }
    const int64_t n = view_len(in);
    // assert!(!(ddof < zero || ddof > n)) in the element type; NaN passes (impl_numeric.rs:185-190)
    const bool f32 = in->dtype == NDB_F32;
    const double nn = f32 ? (double)(float)n : (double)n, dd = f32 ? (double)(float)ddof : ddof;
    if (dd < 0.0 || dd > nn) {
        set_error("`ddof` must not be less than zero or greater than the length of the axis");
        return NDB_ERR_INVALID_ARG;
    }
    DeviceGuard g(ctx->device);
    if (n == 0) {  // 0 / (0 - ddof): NaN for ddof == 0 (tests/numeric.rs:175-188)
        const double v = 0.0 / (nn - dd);
        if (f32) *(float *)out_host = (float)v;
        else *(double *)out_host = v;
        return NDB_OK;
    }
    NDB_TRY(ensure_workspace(ctx, 4096));
    void *dev_scalar = (char *)ctx->workspace + ctx->workspace_bytes - 64;
    if (f32) NDB_TRY((reduce_all_typed<float, R_WELFORD>(ctx, in, (float *)dev_scalar, (float)n - (float)ddof, want_std)));
    else NDB_TRY((reduce_all_typed<double, R_WELFORD>(ctx, in, (double *)dev_scalar, (double)n - ddof, want_std)));
    NDB_CUDA(cudaMemcpyAsync(ctx->host_scalar, dev_scalar, dtype_size(in->dtype), cudaMemcpyDeviceToHost, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out_host, ctx->host_scalar, dtype_size(in->dtype));
    return NDB_OK;
}

int ndb_reduce_axis(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof)
{
    clear_error();
    if (redop < 0 || redop >= NDB_RED__COUNT) {
        set_error("reduce_axis: unknown redop %d", redop);
        return NDB_ERR_UNSUPPORTED;
    }
    return reduce_axis_impl(ctx, redop, axis, in, out, nullptr, ddof);
}

int ndb_reduce_axis_peer(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof, int64_t n_global)
{
    clear_error();
    if (redop < 0 || redop >= NDB_RED__COUNT) {
        set_error("reduce_axis_peer: unknown redop %d", redop);
        return NDB_ERR_UNSUPPORTED;
    }
    if (!ctx || ctx->peer_world <= 0 || !ctx->peer_attached) {
        set_error("reduce_axis_peer: no peer window attached to this context");
        return NDB_ERR_INVALID_ARG;
    }
    return reduce_axis_impl(ctx, redop, axis, in, out, nullptr, ddof, 1, n_global);
}

int ndb_reduce_axis_peer_ok(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof, int64_t n_global)
{
    clear_error();
    if (redop < 0 || redop >= NDB_RED__COUNT || !ctx || ctx->peer_world <= 0 || !ctx->peer_attached) return NDB_ERR_UNSUPPORTED;
    return reduce_axis_impl(ctx, redop, axis, in, out, nullptr, ddof, 2, n_global);
}

int ndb_welford_axis(ndb_ctx *ctx, int axis, const ndb_view *in, const ndb_view *mean_out, const ndb_view *m2_out)
{
    clear_error();
    if (!m2_out) return NDB_ERR_INVALID_ARG;
    return reduce_axis_impl(ctx, -1, axis, in, mean_out, m2_out, 0.0);
}

int ndb_welford_merge(ndb_ctx *ctx, int n_shards, const int64_t *counts, const ndb_view *means, const ndb_view *m2s,
                      double ddof, int want_std, const ndb_view *out)
{
    clear_error();
    if (!ctx || n_shards < 1 || n_shards > 64 || !counts || !means || !m2s || !out) return NDB_ERR_INVALID_ARG;
    // shards and the output are dense C-order arrays of identical shape (they are library-allocated partials)
    const int64_t n_out = view_len(out);
    int64_t total = 0;
    for (int k = 0; k < n_shards; ++k) {
        if (view_len(&means[k]) != n_out || view_len(&m2s[k]) != n_out || means[k].dtype != out->dtype) {
            set_error("welford_merge: shard %d does not match the output", k);
            return NDB_ERR_INCOMPATIBLE_SHAPE;
        }
        if (!is_layout_c(&means[k]) || !is_layout_c(&m2s[k])) {
            set_error("welford_merge: shards must be standard-layout");
            return NDB_ERR_INCOMPATIBLE_LAYOUT;
        }
        total += counts[k];
    }
    if (!is_layout_c(out)) {
        set_error("welford_merge: output must be standard-layout");
        return NDB_ERR_INCOMPATIBLE_LAYOUT;
    }
    if (!dtype_is_float(out->dtype)) return NDB_ERR_UNSUPPORTED;
    if (ddof < 0.0 || ddof > (double)total) {
        set_error("`ddof` must not be less than zero or greater than the length of the axis");
        return NDB_ERR_INVALID_ARG;
    }
    if (n_out == 0) return NDB_OK;
    DeviceGuard g(ctx->device);
    // tables of shard pointers + counts go through the workspace tail
    const size_t tbl = (size_t)n_shards * (sizeof(int64_t) + 2 * sizeof(void *));
    NDB_TRY(ensure_workspace(ctx, 8192));
    char *dev_tbl = (char *)ctx->workspace + ctx->workspace_bytes - 4096;
    std::vector<char> host(tbl);
    int64_t *hc = (int64_t *)host.data();
    const void **hm = (const void **)(host.data() + n_shards * sizeof(int64_t));
    const void **h2 = hm + n_shards;
    for (int k = 0; k < n_shards; ++k) {
        hc[k] = counts[k];
        hm[k] = means[k].ptr;
        h2[k] = m2s[k].ptr;
    }
    NDB_CUDA(cudaMemcpyAsync(dev_tbl, host.data(), tbl, cudaMemcpyHostToDevice, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` dies at scope exit
    const int64_t *dc = (const int64_t *)dev_tbl;
    const unsigned blocks = (unsigned)((n_out + 255) / 256);
    if (out->dtype == NDB_F32) {
        const float dof = (float)total - (float)ddof;
        launch_pdl(welford_merge_kernel<float>, blocks, 256, 0, ctx->stream, 
            n_shards, dc, (const float *const *)(dev_tbl + n_shards * sizeof(int64_t)),
            (const float *const *)(dev_tbl + n_shards * (sizeof(int64_t) + sizeof(void *))), n_out, dof, want_std,
            (float *)out->ptr);
    } else {
        const double dof = (double)total - ddof;
        launch_pdl(welford_merge_kernel<double>, blocks, 256, 0, ctx->stream, 
            n_shards, dc, (const double *const *)(dev_tbl + n_shards * sizeof(int64_t)),
            (const double *const *)(dev_tbl + n_shards * (sizeof(int64_t) + sizeof(void *))), n_out, dof, want_std,
            (double *)out->ptr);
    }
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

int ndb_dot(ndb_ctx *ctx, const ndb_view *x, const ndb_view *y, void *out_host)
{
    clear_error();
    if (!ctx || !out_host) return NDB_ERR_INVALID_ARG;
    NDB_TRY(validate_view(x, "dot lhs"));
    NDB_TRY(validate_view(y, "dot rhs"));
    if (x->ndim != 1 || y->ndim != 1 || x->dim[0] != y->dim[0] || x->dtype != y->dtype) {
        set_error("dot: operands must be 1-d, of equal length and element type");
        return NDB_ERR_INCOMPATIBLE_SHAPE;
    }
    DeviceGuard g(ctx->device);
    const int64_t n = x->dim[0];
    const int max_blocks = ctx->sm_count * 8;
    NDB_TRY(ensure_workspace(ctx, (size_t)max_blocks * 32 + 4096));
    void *dev_scalar = (char *)ctx->workspace + ctx->workspace_bytes - 64;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, (n + 1023) / 1024));
    const int64_t sx = n > 1 ? x->stride[0] : 1, sy = n > 1 ? y->stride[0] : 1;
    const bool dense = sx == 1 && sy == 1 && ((uintptr_t)x->ptr % 16) == 0 && ((uintptr_t)y->ptr % 16) == 0 && n >= 64;
#define NDB_DOT(T)                                                                                            \
    if (dense) launch_pdl(dot_partials_vec<T>, blocks, 256, 0, ctx->stream, (const T *)x->ptr, (const T *)y->ptr, n, ctx->workspace); \
    else launch_pdl(dot_partials<T>, blocks, 256, 0, ctx->stream, (const T *)x->ptr, sx, (const T *)y->ptr, sy, n,    \
                                                     ctx->workspace);                                         \
    NDB_CHECK_LAUNCH();                                                                                       \
    launch_pdl(red_all_finish<T, R_SUM>, 1, 256, 0, ctx->stream, ctx->workspace, blocks, (T *)dev_scalar, T(1), 0);   \
    NDB_CHECK_LAUNCH();
    switch (x->dtype) {
    case NDB_F32: { NDB_DOT(float) } break;
    case NDB_F64: { NDB_DOT(double) } break;
    case NDB_I32: { NDB_DOT(int32_t) } break;
    case NDB_I64: { NDB_DOT(int64_t) } break;
    default: return NDB_ERR_UNSUPPORTED;
    }
#undef NDB_DOT
    NDB_CUDA(cudaMemcpyAsync(ctx->host_scalar, dev_scalar, dtype_size(x->dtype), cudaMemcpyDeviceToHost, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out_host, ctx->host_scalar, dtype_size(x->dtype));
    return NDB_OK;
}

}  // extern "C"
