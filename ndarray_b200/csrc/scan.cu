// accumulate_axis_inplace (src/impl_methods.rs:3243) / cumprod (src/numeric/impl_numeric.rs:125) as a scan:
//   a[.., i, ..] = op(a[.., i-1, ..], a[.., i, ..])   for i = 1 .. n-1 along `axis`, in place.
// ndarray walks the axis sequentially with one elementwise pass per step.  Here:
//   scan_cols   a kept axis is the dense one: a thread owns one lane and walks the scanned axis sequentially (exactly
//               ndarray's association, so results are bit-identical for floats too); neighbouring threads touch
//               neighbouring addresses, every step is coalesced.
//   scan_lanes  the scanned axis is the dense one: a warp owns a lane and scans it 32 elements at a time with shuffles
//               plus a running carry (coalesced; the association differs, so floats agree to rounding, integers exactly).
//   scan_chunk_* a long strided axis with few lanes (n >= 512): the axis is cut into chunks scanned by separate
//               threads; this re-associates float sums too (tolerance-class), integers stay exact.
// Exactness by path: scan_cols = ndarray's association (floats bit-identical); scan_lanes(_vec) and scan_chunk_* =
// re-associated (floats to rounding).  No path folds an identity into element 0, so a leading -0.0 is preserved.
#include <algorithm>

#include "zip_kernels.cuh"
#include "zip_plan.h"

namespace ndb {

namespace {

struct ScanArgs {
    char *base;
    int n_kept;
    int64_t kdim[NDB_MAX_NDIM], kstride[NDB_MAX_NDIM];
    int64_t n_lanes;
    int64_t n, s;  // scanned axis: length and stride (elements)
};

template <typename T, bool MUL>
__device__ __forceinline__ T comb(T a, T b)
{
    return MUL ? w_mul(a, b) : w_add(a, b);
}

__device__ __forceinline__ int64_t lane_offset(const ScanArgs &a, int64_t lin)
{
    int64_t off = 0;
    for (int d = a.n_kept - 1; d >= 0; --d) {
        const int64_t q = lin / a.kdim[d];
        off += (lin - q * a.kdim[d]) * a.kstride[d];
        lin = q;
    }
    return off;
}

template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_cols(ScanArgs a)
{
    pdl_enter();
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= a.n_lanes) return;
    T *p = reinterpret_cast<T *>(a.base) + lane_offset(a, lane);
    T acc = p[0];
    int64_t i = 1;
    for (; i + 3 < a.n; i += 4) {  // four independent loads in flight, combined in order (a generic x[U] form with U = 4 / 8 / 16
                                   // measured 4.3 / 5.2 / 5.6 TB/s against 5.75 for this one)
        const T x0 = p[i * a.s], x1 = p[(i + 1) * a.s], x2 = p[(i + 2) * a.s], x3 = p[(i + 3) * a.s];
        acc = comb<T, MUL>(acc, x0);
        p[i * a.s] = acc;
        acc = comb<T, MUL>(acc, x1);
        p[(i + 1) * a.s] = acc;
        acc = comb<T, MUL>(acc, x2);
        p[(i + 2) * a.s] = acc;
        acc = comb<T, MUL>(acc, x3);
        p[(i + 3) * a.s] = acc;
    }
    for (; i < a.n; ++i) {
        acc = comb<T, MUL>(acc, p[i * a.s]);
        p[i * a.s] = acc;
    }
}

template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_lanes(ScanArgs a)
{
    pdl_enter();
    const int lane_id = threadIdx.x & 31;
    const int64_t lane = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (lane >= a.n_lanes) return;
    T *p = reinterpret_cast<T *>(a.base) + lane_offset(a, lane);
    T carry = MUL ? T(1) : T(0);
    for (int64_t i0 = 0; i0 < a.n; i0 += 32) {
        const int64_t i = i0 + lane_id;
        const bool live = i < a.n;
        T v = live ? p[i * a.s] : (MUL ? T(1) : T(0));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, v, d);
            if (lane_id >= d) v = comb<T, MUL>(up, v);
        }
        // no identity is ever folded in: element 0 stays as it is (-0.0 stays -0.0, impl_methods.rs:3243 starts at 1)
        if (i0 > 0) v = comb<T, MUL>(carry, v);
        if (live) p[i * a.s] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// ---- few lanes, long strided axis: chunk the scanned axis ---------------------------------------------------------
// scan_cols gives one thread per lane; 8192 lanes of 8192 elements are 32 blocks on 148 SMs (0.86 TB/s).  Three small
// kernels instead: per (chunk, lane) totals; an exclusive scan of the totals along the chunks (one thread per lane);
// the local scans again with their carry-in.  The data is read twice, but by ~300 K threads.
template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_chunk_totals(ScanArgs a, T *totals, int chunks, int64_t chunk_len)
{
    pdl_enter();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.n_lanes * chunks) return;
    const int64_t lane = idx % a.n_lanes, chunk = idx / a.n_lanes;  // neighbouring threads -> neighbouring lanes
    const T *p = reinterpret_cast<const T *>(a.base) + lane_offset(a, lane);
    const int64_t c0 = chunk * chunk_len, c1 = c0 + chunk_len < a.n ? c0 + chunk_len : a.n;
    T acc = p[c0 * a.s];
    for (int64_t i = c0 + 1; i < c1; ++i) acc = comb<T, MUL>(acc, p[i * a.s]);
    totals[idx] = acc;
}

template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_chunk_carries(T *totals, int64_t n_lanes, int chunks)
{
    pdl_enter();
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n_lanes) return;
    T running = totals[lane];  // carry into chunk 1 = total of chunk 0; chunk 0 itself takes no carry
    for (int c = 1; c < chunks; ++c) {
        const T t = totals[(int64_t)c * n_lanes + lane];
        totals[(int64_t)c * n_lanes + lane] = running;
        running = comb<T, MUL>(running, t);
    }
}

template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_chunk_apply(ScanArgs a, const T *carries, int chunks, int64_t chunk_len)
{
    pdl_enter();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.n_lanes * chunks) return;
    const int64_t lane = idx % a.n_lanes, chunk = idx / a.n_lanes;
    T *p = reinterpret_cast<T *>(a.base) + lane_offset(a, lane);
    const int64_t c0 = chunk * chunk_len, c1 = c0 + chunk_len < a.n ? c0 + chunk_len : a.n;
    int64_t i = c0;
    T acc;
    if (chunk == 0) {
        acc = p[0];  // the first element is left as it is (no identity is folded in: -0.0 stays -0.0)
        i = 1;
    } else {
        acc = carries[idx];
    }
    for (; i < c1; ++i) {
        acc = comb<T, MUL>(acc, p[i * a.s]);
        p[i * a.s] = acc;
    }
}

// scan_lanes with 128-bit accesses: a thread scans one vector locally, the warp scans the 32 vector totals with shuffles,
// so a 32*N-element step costs one shuffle scan instead of N of them (unit-stride, 16-byte aligned lanes).  The steps of a
// lane are sequential (each needs the carry of the one before), so the LOADS of U steps are issued together before the
// first of them is scanned: a warp has U vectors in flight instead of one.  Measured: f32 512^3 along the dense axis 6.08 -> 6.62 TB/s with U = 4 (6.54 with 8); f64 8192^2
// 5.29 -> 5.50 with U = 4, 5.89 with U = 8.
template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_lanes_vec(ScanArgs a)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    constexpr int U = sizeof(T) == 8 ? 8 : 4;
    const int lane_id = threadIdx.x & 31;
    const int64_t lane = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (lane >= a.n_lanes) return;
    T *p = reinterpret_cast<T *>(a.base) + lane_offset(a, lane);
    const T ident = MUL ? T(1) : T(0);
    T carry = ident;
    for (int64_t s0 = 0; s0 < a.n; s0 += (int64_t)U * 32 * N) {
        V vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = s0 + (int64_t)u * 32 * N + (int64_t)lane_id * N;
            if (i + N <= a.n) {
                vv[u] = *reinterpret_cast<const V *>(p + i);
            } else {
#pragma unroll
                for (int e = 0; e < N; ++e) vv[u].v[e] = (i + e < a.n) ? p[i + e] : ident;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i0 = s0 + (int64_t)u * 32 * N;
            if (i0 >= a.n) break;  // warp-uniform
            const int64_t i = i0 + (int64_t)lane_id * N;
            const bool whole = i + N <= a.n;
            V v = vv[u];
#pragma unroll
            for (int e = 1; e < N; ++e) v.v[e] = comb<T, MUL>(v.v[e - 1], v.v[e]);
            T inc = v.v[N - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T up = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane_id >= d) inc = comb<T, MUL>(up, inc);
            }
            const T excl = __shfl_up_sync(0xffffffffu, inc, 1);
            // what precedes this thread's vector: nothing (first vector of the lane), the carry alone (thread 0 of a later
            // step), the warp prefix alone (first step) or both - an identity is never folded in, so -0.0 survives
            const bool first = i0 == 0;
            const T off = lane_id == 0 ? carry : (first ? excl : comb<T, MUL>(carry, excl));
            if (!(first && lane_id == 0)) {
#pragma unroll
                for (int e = 0; e < N; ++e) v.v[e] = comb<T, MUL>(off, v.v[e]);
            }
            if (whole) {
                *reinterpret_cast<V *>(p + i) = v;
            } else {
#pragma unroll
                for (int e = 0; e < N; ++e)
                    if (i + e < a.n) p[i + e] = v.v[e];
            }
            const T step_total = __shfl_sync(0xffffffffu, inc, 31);
            carry = first ? step_total : comb<T, MUL>(carry, step_total);
        }
    }
}

// ---- few lanes, long strided axis, ONE pass: chained scan ---------------------------------------------------------------
// The chunked form above reads the data twice (3.5 TB/s for f64 8192^2 down the rows).  Here a block owns 64 rows x
// (32 x one 128-bit vector) columns, keeps them in registers, scans them locally and then waits for ONE value per column
// from the block above it (the inclusive prefix of everything before this chunk), adds its own total and passes the result
// on before it even starts writing.  A value and the launch's epoch travel in one 8- or 16-byte packet written with a
// single store, so there is no separate flag, no fence and the packets never need clearing (epochs only grow).  Blocks take
// their position from a ticket counter, chunk-major, so a block's predecessor has always started: no deadlock however the
// hardware orders blocks.  The chain is strictly sequential over the chunks (prefix_c = prefix_{c-1} o total_c), so the
// result does not depend on timing; the data is read once and written once.
template <typename T>
struct ScanPacket;
template <>
struct alignas(8) ScanPacket<float> { float v; uint32_t e; };
template <>
struct alignas(8) ScanPacket<int32_t> { int32_t v; uint32_t e; };
template <>
struct alignas(16) ScanPacket<double> { double v; uint64_t e; };
template <>
struct alignas(16) ScanPacket<int64_t> { int64_t v; uint64_t e; };

template <typename T>
__device__ __forceinline__ void packet_store(ScanPacket<T> *p, T v, uint32_t epoch)
{
    if constexpr (sizeof(T) == 4) {
        const uint32_t bits = *reinterpret_cast<const uint32_t *>(&v);
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(epoch) : "memory");
    } else {
        const unsigned long long bits = *reinterpret_cast<const unsigned long long *>(&v);
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(bits), "l"((unsigned long long)epoch) : "memory");
    }
}

// spins until the packet carries this launch's epoch
template <typename T>
__device__ __forceinline__ T packet_wait(const ScanPacket<T> *p, uint32_t epoch)
{
    if constexpr (sizeof(T) == 4) {
        uint32_t bits, e;
        do {
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(e) : "l"(p) : "memory");
        } while (e != epoch);
        return *reinterpret_cast<const T *>(&bits);
    } else {
        unsigned long long bits, e;
        do {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bits), "=l"(e) : "l"(p) : "memory");
        } while (e != (unsigned long long)epoch);
        return *reinterpret_cast<const T *>(&bits);
    }
}

#ifndef NDB_SCAN_CHAIN_R
#define NDB_SCAN_CHAIN_R 16  // measured f64 8192^2 down the rows: R = 4 / 8 / 16 -> 3.6 / 5.1 / 6.1 TB/s (the chain has 8192 / (8 R) hops)
#endif
constexpr int SCAN_CHAIN_R = NDB_SCAN_CHAIN_R;        // rows per thread
constexpr int SCAN_CHAIN_ROWS = 8 * SCAN_CHAIN_R;    // rows per block

template <typename T, bool MUL>
__global__ void __launch_bounds__(256) scan_cols_chained(ScanArgs a, ScanPacket<T> *packets, unsigned *ticket, uint32_t epoch, int tiles,
                                                         int chunks)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int W = V::N;
    constexpr int R = SCAN_CHAIN_R;
    __shared__ T sm_tot[8][32 * W];  // per row-group totals, turned into exclusive offsets in place
    __shared__ T sm_pred[32 * W];    // inclusive prefix of everything above this chunk
    __shared__ unsigned sm_bid;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (tx == 0 && ty == 0) sm_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = sm_bid;
    const int chunk = (int)(bid / (unsigned)tiles), tile = (int)(bid % (unsigned)tiles);
    T *p = reinterpret_cast<T *>(a.base) + lane_offset(a, (int64_t)tile * (32 * W) + (int64_t)tx * W);
    const int64_t r0 = (int64_t)chunk * SCAN_CHAIN_ROWS + (int64_t)ty * R;
    V v[R];
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = *reinterpret_cast<const V *>(p + (r0 + k) * a.s);
#pragma unroll
    for (int k = 1; k < R; ++k)
#pragma unroll
        for (int e = 0; e < W; ++e) v[k].v[e] = comb<T, MUL>(v[k - 1].v[e], v[k].v[e]);
#pragma unroll
    for (int e = 0; e < W; ++e) sm_tot[ty][tx * W + e] = v[R - 1].v[e];
    __syncthreads();
    if (ty == 0) {
        ScanPacket<T> *mine = packets + ((size_t)chunk * tiles + tile) * (32 * W) + tx * W;
        const ScanPacket<T> *above = mine - (size_t)tiles * (32 * W);
#pragma unroll
        for (int e = 0; e < W; ++e) {
            const int c = tx * W + e;
            T running = sm_tot[0][c];
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const T t = sm_tot[j][c];
                sm_tot[j][c] = running;  // what precedes row-group j inside this chunk
                running = comb<T, MUL>(running, t);
            }
            if (chunk > 0) {
                const T pred = packet_wait<T>(above + e, epoch);
                packet_store<T>(mine + e, comb<T, MUL>(pred, running), epoch);  // pass the chain on first
                sm_pred[c] = pred;
            } else {
                packet_store<T>(mine + e, running, epoch);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < W; ++e) {
        const int c = tx * W + e;
        // nothing is folded into the very first rows (element 0 stays as it is, -0.0 included)
        if (chunk > 0 || ty > 0) {
            const T off = ty == 0 ? sm_pred[c] : (chunk > 0 ? comb<T, MUL>(sm_pred[c], sm_tot[ty][c]) : sm_tot[ty][c]);
#pragma unroll
            for (int k = 0; k < R; ++k) v[k].v[e] = comb<T, MUL>(off, v[k].v[e]);
        }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) *reinterpret_cast<V *>(p + (r0 + k) * a.s) = v[k];
    if (tx == 0 && ty == 0 && bid == (unsigned)tiles * (unsigned)chunks - 1u) *ticket = 0;  // the last ticket: ready for the next launch
}

// the chained kernel's geometry: whole 64-row chunks, whole 32-vector column tiles, every row vector 16-byte aligned
template <typename T>
static bool chain_eligible(const ScanArgs &a)
{
    constexpr int64_t W = 16 / (int64_t)sizeof(T);
    if (a.n_kept < 1 || a.n % SCAN_CHAIN_ROWS != 0 || ((uintptr_t)a.base % 16) != 0 || a.s % W != 0) return false;
    const int last = a.n_kept - 1;
    if (a.kstride[last] != 1 || a.kdim[last] % (32 * W) != 0) return false;
    for (int d = 0; d < last; ++d)
        if (a.kdim[d] > 1 && a.kstride[d] % W != 0) return false;
    return a.n_lanes / (32 * W) * (a.n / SCAN_CHAIN_ROWS) < ((int64_t)1 << 31);
}

// ticket + packets of the chained scan: a dedicated allocation (never used as scratch by anything else, zeroed when it is
// created), so a packet can only ever carry an epoch that a chained scan on this context wrote
static int ensure_scan_chain(ndb_ctx *ctx, size_t bytes)
{
    if (ctx->scan_chain_bytes >= bytes) return NDB_OK;
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->scan_chain) NDB_CUDA(cudaFree(ctx->scan_chain));
    ctx->scan_chain = nullptr;
    ctx->scan_chain_bytes = 0;
    NDB_CUDA(cudaMalloc(&ctx->scan_chain, bytes));
    NDB_CUDA(cudaMemsetAsync(ctx->scan_chain, 0, bytes, ctx->stream));
    ctx->scan_chain_bytes = bytes;
    ctx->scan_epoch = 0;
    return NDB_OK;
}

// The chained kernel's epoch is a launch ARGUMENT: a captured graph would replay it and find last replay's packets already
// "valid".  While the stream is capturing (stream_is_capturing, ctx.cu) the chunked two-pass form is used instead.
template <typename T>
int scan_dispatch(ndb_ctx *ctx, int op, const ScanArgs &a, bool lanes)
{
    const bool mul = op == NDB_OP_MUL;
    if (lanes) {
        const unsigned blocks = (unsigned)((a.n_lanes + 7) / 8);
        constexpr int64_t VN = 16 / sizeof(T);
        bool vec = a.s == 1 && ((uintptr_t)a.base % 16) == 0 && a.n >= 2 * VN && !getenv("NDB_SCAN_SCALAR");
        for (int d = 0; d < a.n_kept && vec; ++d) vec = a.kdim[d] == 1 || a.kstride[d] % VN == 0;
        if (vec) {
            if (mul) launch_pdl(scan_lanes_vec<T, true>, blocks, 256, 0, ctx->stream, a);
            else launch_pdl(scan_lanes_vec<T, false>, blocks, 256, 0, ctx->stream, a);
        } else if (mul) launch_pdl(scan_lanes<T, true>, blocks, 256, 0, ctx->stream, a);
        else launch_pdl(scan_lanes<T, false>, blocks, 256, 0, ctx->stream, a);
    } else if (a.n >= 512 && a.n_lanes < (int64_t)ctx->sm_count * 512 && !getenv("NDB_SCAN_NO_CHUNKS") && !getenv("NDB_SCAN_NO_CHAIN") &&
               chain_eligible<T>(a) && !stream_is_capturing(ctx)) {
        constexpr int W = 16 / (int)sizeof(T);
        const int tiles = (int)(a.n_lanes / (32 * W)), chunks = (int)(a.n / SCAN_CHAIN_ROWS);
        const size_t need = 256 + (size_t)tiles * chunks * 32 * W * sizeof(ScanPacket<T>);
        NDB_TRY(ensure_scan_chain(ctx, need));
        unsigned *ticket = reinterpret_cast<unsigned *>(ctx->scan_chain);
        ScanPacket<T> *packets = reinterpret_cast<ScanPacket<T> *>((char *)ctx->scan_chain + 256);
        if (ctx->scan_epoch == UINT32_MAX) {  // epochs only grow: start over from clean packets before the counter wraps
            NDB_CUDA(cudaMemsetAsync(ctx->scan_chain, 0, ctx->scan_chain_bytes, ctx->stream));
            ctx->scan_epoch = 0;
        }
        const uint32_t epoch = ++ctx->scan_epoch;
        const dim3 grid((unsigned)tiles * (unsigned)chunks), block(32, 8, 1);
        if (mul) launch_pdl(scan_cols_chained<T, true>, grid, block, 0, ctx->stream, a, packets, ticket, epoch, tiles, chunks);
        else launch_pdl(scan_cols_chained<T, false>, grid, block, 0, ctx->stream, a, packets, ticket, epoch, tiles, chunks);
    } else if (a.n >= 512 && a.n_lanes < (int64_t)ctx->sm_count * 512 && !getenv("NDB_SCAN_NO_CHUNKS")) {
        const int64_t target = (int64_t)ctx->sm_count * 2048;  // threads that fill the machine
        int chunks = (int)std::min<int64_t>((target + a.n_lanes - 1) / a.n_lanes, a.n / 64);
        chunks = std::max(2, std::min(chunks, 4096));
        const int64_t chunk_len = (a.n + chunks - 1) / chunks;
        chunks = (int)((a.n + chunk_len - 1) / chunk_len);  // no empty chunk
        NDB_TRY(ensure_workspace(ctx, (size_t)chunks * (size_t)a.n_lanes * sizeof(T)));
        T *totals = reinterpret_cast<T *>(ctx->workspace);
        const unsigned blocks = (unsigned)((a.n_lanes * chunks + 255) / 256);
        const unsigned lane_blocks = (unsigned)((a.n_lanes + 255) / 256);
        if (mul) {
            launch_pdl(scan_chunk_totals<T, true>, blocks, 256, 0, ctx->stream, a, totals, chunks, chunk_len);
            launch_pdl(scan_chunk_carries<T, true>, lane_blocks, 256, 0, ctx->stream, totals, a.n_lanes, chunks);
            launch_pdl(scan_chunk_apply<T, true>, blocks, 256, 0, ctx->stream, a, totals, chunks, chunk_len);
        } else {
            launch_pdl(scan_chunk_totals<T, false>, blocks, 256, 0, ctx->stream, a, totals, chunks, chunk_len);
            launch_pdl(scan_chunk_carries<T, false>, lane_blocks, 256, 0, ctx->stream, totals, a.n_lanes, chunks);
            launch_pdl(scan_chunk_apply<T, false>, blocks, 256, 0, ctx->stream, a, totals, chunks, chunk_len);
        }
    } else {
        const unsigned blocks = (unsigned)((a.n_lanes + 255) / 256);
        if (mul) launch_pdl(scan_cols<T, true>, blocks, 256, 0, ctx->stream, a);
        else launch_pdl(scan_cols<T, false>, blocks, 256, 0, ctx->stream, a);
    }
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_accumulate_axis(ndb_ctx *ctx, int op, int axis, const ndb_view *v)
{
    clear_error();
    if (!ctx) {
        set_error("accumulate_axis: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    NDB_TRY(validate_view(v, "accumulate_axis"));
    if (op != NDB_OP_ADD && op != NDB_OP_MUL) {
        set_error("accumulate_axis: op must be NDB_OP_ADD (running sum) or NDB_OP_MUL (cumprod)");
        return NDB_ERR_UNSUPPORTED;
    }
    if (axis < 0 || axis >= v->ndim) {
        set_error("axis %d is out of bounds for array of dimension %d", axis, v->ndim);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    const int64_t n = v->dim[axis];
    for (int i = 0; i < v->ndim; ++i)
        if (v->dim[i] == 0) return NDB_OK;
    if (n <= 1) return NDB_OK;
    {
        // the WHOLE view is written in place: a zero or aliasing stride on the scanned axis would race
        ndb_view whole[2] = {*v, *v};
        ZipPlan zw;
        NDB_TRY(make_zip_plan(NDB_OP_COPY, 2, whole, &zw));
    }
    // collapse the kept axes of the (written) view with the Zip planner: checks self-aliasing, flips, merges
    ndb_view kept = *v;
    for (int i = axis; i + 1 < v->ndim; ++i) {
        kept.dim[i] = v->dim[i + 1];
        kept.stride[i] = v->stride[i + 1];
    }
    kept.ndim = v->ndim - 1;
    ndb_view parts[2] = {kept, kept};
    ZipPlan z;
    NDB_TRY(make_zip_plan(NDB_OP_COPY, 2, parts, &z));
    ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.base = z.ptr[0];
    a.n_kept = z.ndim;
    a.n_lanes = z.total;
    int64_t min_kept = INT64_MAX;
    for (int d = 0; d < z.ndim; ++d) {
        a.kdim[d] = z.dim[d];
        a.kstride[d] = z.stride[0][d];
        if (z.dim[d] > 1) min_kept = std::min<int64_t>(min_kept, llabs(z.stride[0][d]));
    }
    a.n = n;
    a.s = v->stride[axis];
    const bool lanes = a.n_lanes == 1 || llabs(a.s) < min_kept;
    DeviceGuard g(ctx->device);
    switch (v->dtype) {
    case NDB_F32: return scan_dispatch<float>(ctx, op, a, lanes);
    case NDB_F64: return scan_dispatch<double>(ctx, op, a, lanes);
    case NDB_I32: return scan_dispatch<int32_t>(ctx, op, a, lanes);
    case NDB_I64: return scan_dispatch<int64_t>(ctx, op, a, lanes);
    }
    return NDB_ERR_UNSUPPORTED;
}
