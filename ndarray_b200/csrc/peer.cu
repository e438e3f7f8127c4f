// Peer windows (CUDA IPC) and the small-vector exchanges built on them (peer.cuh): all-reduce and all-gather of the
// partial results the sharded reductions produce: sum()/dot -> 1 element, sum_axis(Axis(0)) -> [n_out],
// var_axis(Axis(0)) -> (mean, M2) vectors.  One launch per exchange, at most one block per SM so that every block of
// every rank is resident while it waits for the others.
#include "peer.cuh"
#include "zip_kernels.cuh"

namespace ndb {

namespace {

template <typename T, int REDOP>
__device__ __forceinline__ T combine(T a, T b)
{
    if constexpr (REDOP == NDB_RED_SUM) return w_add(a, b);
    else if constexpr (REDOP == NDB_RED_PROD) return w_mul(a, b);
    else if constexpr (REDOP == NDB_RED_MIN) return w_min(a, b);
    else return w_max(a, b);
}

// REDOP < 0: all-gather (out is [world][n]); else all-reduce in rank order (out is [n], may alias local)
template <typename T, int REDOP>
__global__ void __launch_bounds__(256) peer_exchange(PeerArgs p, const T *local, T *out, int64_t n)
{
    pdl_wait_only();
    constexpr int VEC = 16 / sizeof(T);
    int64_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + VEC - 1) / VEC * VEC;  // slices start on 16-byte boundaries
    const int64_t lo = (int64_t)blockIdx.x * per < n ? (int64_t)blockIdx.x * per : n;
    const int64_t hi = lo + per < n ? lo + per : n;
    const bool vec = (reinterpret_cast<uintptr_t>(local) & 15) == 0;  // window slots are 256-byte aligned
    for (int r = 0; r < p.world; ++r) {
        T *dst = reinterpret_cast<T *>(peer_slot(p, r, p.rank));
        if (vec) {
            const int64_t nv = (hi - lo) / VEC;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(local + lo);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst + lo);
            for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = s4[i];
            for (int64_t i = lo + nv * VEC + threadIdx.x; i < hi; i += blockDim.x) dst[i] = local[i];
        } else {
            for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = local[i];
        }
    }
    peer_arrive(p);
    peer_wait(p);
    if constexpr (REDOP < 0) {
        for (int r = 0; r < p.world; ++r) {
            const T *src = reinterpret_cast<const T *>(peer_slot(p, p.rank, r));
            for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) out[(int64_t)r * n + i] = __ldcg(src + i);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            T acc = __ldcg(reinterpret_cast<const T *>(peer_slot(p, p.rank, 0)) + i);
            for (int r = 1; r < p.world; ++r) acc = combine<T, REDOP>(acc, __ldcg(reinterpret_cast<const T *>(peer_slot(p, p.rank, r)) + i));
            out[i] = acc;
        }
    }
}

template <typename T>
int exchange_typed(ndb_ctx *ctx, int redop, const PeerArgs &p, int blocks, const void *local, void *out, int64_t n)
{
    const T *l = (const T *)local;
    T *o = (T *)out;
    switch (redop) {
    case -1: peer_exchange<T, -1><<<blocks, 256, 0, ctx->stream>>>(p, l, o, n); break;
    case NDB_RED_SUM: peer_exchange<T, NDB_RED_SUM><<<blocks, 256, 0, ctx->stream>>>(p, l, o, n); break;
    case NDB_RED_PROD: peer_exchange<T, NDB_RED_PROD><<<blocks, 256, 0, ctx->stream>>>(p, l, o, n); break;
    case NDB_RED_MIN: peer_exchange<T, NDB_RED_MIN><<<blocks, 256, 0, ctx->stream>>>(p, l, o, n); break;
    case NDB_RED_MAX: peer_exchange<T, NDB_RED_MAX><<<blocks, 256, 0, ctx->stream>>>(p, l, o, n); break;
    default:
        set_error("peer exchange: redop %d is not sum/product/min/max", redop);
        return NDB_ERR_UNSUPPORTED;
    }
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

int exchange(ndb_ctx *ctx, int redop, int dtype, const void *local, void *out, int64_t n)
{
    if (n <= 0) return NDB_OK;
    const size_t bytes = (size_t)n * dtype_size(dtype);
    // every rank derives the same block count from n alone (the arrival target depends on it)
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (int64_t)((bytes + 2047) / 2048)));
    PeerArgs p;
    NDB_TRY(peer_begin(ctx, blocks, bytes, &p));
    DeviceGuard g(ctx->device);
    switch (dtype) {
    case NDB_F32: return exchange_typed<float>(ctx, redop, p, blocks, local, out, n);
    case NDB_F64: return exchange_typed<double>(ctx, redop, p, blocks, local, out, n);
    case NDB_I32: return exchange_typed<int32_t>(ctx, redop, p, blocks, local, out, n);
    case NDB_I64: return exchange_typed<int64_t>(ctx, redop, p, blocks, local, out, n);
    }
    return NDB_ERR_UNSUPPORTED;
}

}  // namespace

// Reserves the next exchange on this context: which slot set it uses and the counter value that means "all arrived".
int peer_begin(ndb_ctx *ctx, int blocks, size_t bytes, PeerArgs *out)
{
    if (ctx->peer_world <= 0 || !ctx->peer_attached) {
        set_error("peer exchange: no peer window attached to this context (ndb_peer_window_create / _attach)");
        return NDB_ERR_INVALID_ARG;
    }
    if (bytes > ctx->peer_slot_bytes) {
        set_error("peer exchange: %zu bytes exceed the window's slot size %zu", bytes, ctx->peer_slot_bytes);
        return NDB_ERR_OVERFLOW;
    }
    memset(out, 0, sizeof(*out));
    for (int r = 0; r < ctx->peer_world; ++r) out->win[r] = (char *)ctx->peer_win[r];
    out->rank = ctx->peer_rank;
    out->world = ctx->peer_world;
    ctx->peer_expected += (unsigned long long)ctx->peer_world * (unsigned long long)blocks;
    out->target = ctx->peer_expected;
    out->slot_bytes = ctx->peer_slot_bytes;
    out->set_off = PEER_HEADER + (size_t)(ctx->peer_calls & 1u) * (size_t)ctx->peer_world * ctx->peer_slot_bytes;
    ctx->peer_calls++;
    return NDB_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_peer_window_create(ndb_ctx *ctx, int rank, int world, size_t slot_bytes, void *handle_out)
{
    clear_error();
    if (!ctx || !handle_out || world < 1 || world > NDB_MAX_PEERS || rank < 0 || rank >= world || slot_bytes == 0) {
        set_error("peer_window_create: need 0 <= rank < world <= %d and a non-zero slot size", NDB_MAX_PEERS);
        return NDB_ERR_INVALID_ARG;
    }
    if (ctx->peer_world > 0) {
        set_error("peer_window_create: this context already has a window");
        return NDB_ERR_INVALID_ARG;
    }
    DeviceGuard g(ctx->device);
    slot_bytes = (slot_bytes + 255) & ~(size_t)255;
    const size_t total = PEER_HEADER + 2 * (size_t)world * slot_bytes;
    void *base = nullptr;
    NDB_CUDA(cudaMalloc(&base, total));  // cudaMalloc, not the stream-ordered pool: IPC handles need a plain allocation
    NDB_CUDA(cudaMemset(base, 0, total));
    cudaIpcMemHandle_t h;
    NDB_CUDA(cudaIpcGetMemHandle(&h, base));
    static_assert(sizeof(h) == 64, "ndb200.h promises 64-byte handles");
    memcpy(handle_out, &h, sizeof(h));
    NDB_CUDA(cudaMalloc((void **)&ctx->peer_tile_epoch, PEER_MAX_TILES * sizeof(unsigned long long)));
    NDB_CUDA(cudaMemset(ctx->peer_tile_epoch, 0, PEER_MAX_TILES * sizeof(unsigned long long)));
    ctx->peer_rank = rank;
    ctx->peer_world = world;
    ctx->peer_slot_bytes = slot_bytes;
    ctx->peer_win[rank] = base;
    ctx->peer_attached = world == 1;
    ctx->peer_expected = 0;
    ctx->peer_calls = 0;
    return NDB_OK;
}

extern "C" int ndb_peer_window_attach(ndb_ctx *ctx, const void *handles)
{
    clear_error();
    if (!ctx || !handles || ctx->peer_world <= 0) {
        set_error("peer_window_attach: create the local window first");
        return NDB_ERR_INVALID_ARG;
    }
    DeviceGuard g(ctx->device);
    for (int r = 0; r < ctx->peer_world; ++r) {
        if (r == ctx->peer_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        NDB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_win[r] = p;
    }
    ctx->peer_attached = true;
    return NDB_OK;
}

extern "C" int ndb_peer_window_destroy(ndb_ctx *ctx)
{
    clear_error();
    if (!ctx || ctx->peer_world <= 0) return NDB_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->peer_world; ++r) {
        if (!ctx->peer_win[r]) continue;
        if (r == ctx->peer_rank) cudaFree(ctx->peer_win[r]);
        else cudaIpcCloseMemHandle(ctx->peer_win[r]);
        ctx->peer_win[r] = nullptr;
    }
    if (ctx->peer_tile_epoch) cudaFree(ctx->peer_tile_epoch);
    ctx->peer_tile_epoch = nullptr;
    ctx->peer_world = 0;
    ctx->peer_rank = -1;
    ctx->peer_attached = false;
    return NDB_OK;
}

extern "C" int ndb_peer_allreduce(ndb_ctx *ctx, int redop, int dtype, const void *local_dev, void *out_dev, int64_t n)
{
    clear_error();
    if (!ctx || redop < NDB_RED_SUM || redop > NDB_RED_MAX) {
        set_error("peer_allreduce: redop must be sum, product, min or max");
        return NDB_ERR_INVALID_ARG;
    }
    return exchange(ctx, redop, dtype, local_dev, out_dev, n);
}

extern "C" int ndb_peer_allgather(ndb_ctx *ctx, int dtype, const void *local_dev, void *out_dev, int64_t n)
{
    clear_error();
    if (!ctx) return NDB_ERR_INVALID_ARG;
    return exchange(ctx, -1, dtype, local_dev, out_dev, n);
}
