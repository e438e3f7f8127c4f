// Sharded entry points behind the C ABI: what a Rust host calls to run config 3 / 4 / 5 across GPUs without
// any Python - one process per GPU, Axis(0) is the sharded axis.
//
//   ndb_comm_*                 an NCCL communicator the LIBRARY owns (libnccl.so.2 is dlopen'ed on first use, so
//                              libndb200 has no link-time NCCL dependency and single-GPU hosts never load it)
//   ndb_gemm_rowsharded        general_mat_mul with A and C row-sharded, B replicated: ncclBroadcast of B on the context
//                              stream, then the local tensor-core GEMM
//   ndb_reduce_axis_sharded    sum / product / min / max / mean / var / std along the SHARDED axis: the fused one-launch
//                              kernel over peer windows when they are attached and the layout qualifies (all ranks decide
//                              alike, see ndb_reduce_axis_peer), else local partial + ncclAllReduce, or per-shard Welford
//                              states + ncclAllGather + Chan merge in rank order
//   ndb_reduce_all_sharded     global sum()/product()/min/max: peer-fused finish when windows are attached, else NCCL
#include <dlfcn.h>

#include <mutex>

#include "zip_plan.h"

namespace ndb {

namespace {

// the few NCCL declarations we need (ABI-stable since NCCL 2.0)
struct NcclUniqueId {
    char internal[128];
};
using nccl_comm_t = void *;
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_PROD = 1, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = nullptr;
        const char *names[] = {getenv("NDB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return;
#define NDB_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
        NDB_SYM(GetUniqueId, "ncclGetUniqueId");
        NDB_SYM(CommInitRank, "ncclCommInitRank");
        NDB_SYM(CommDestroy, "ncclCommDestroy");
        NDB_SYM(Broadcast, "ncclBroadcast");
        NDB_SYM(AllReduce, "ncclAllReduce");
        NDB_SYM(AllGather, "ncclAllGather");
        NDB_SYM(GetErrorString, "ncclGetErrorString");
#undef NDB_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllReduce && api.AllGather;
    });
    return api;
}

int nccl_fail(int rc, const char *what)
{
    set_error("NCCL error %d (%s) in %s", rc, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?", what);
    return NDB_ERR_CUDA;
}

#define NDB_NCCL(call, what)                  \
    do {                                      \
        int rc_ = (call);                     \
        if (rc_ != 0) return nccl_fail(rc_, what); \
    } while (0)

int nccl_dtype(int dt)
{
    switch (dt) {
    case NDB_F32: return NCCL_FLOAT32;
    case NDB_F64: return NCCL_FLOAT64;
    case NDB_I32: return NCCL_INT32;
    default: return NCCL_INT64;
    }
}

int need_comm(ndb_ctx *ctx, const char *who)
{
    if (!ctx) {
        set_error("%s: null context (no CPU fallback exists)", who);
        return NDB_ERR_INVALID_ARG;
    }
    if (!ctx->nccl_comm) {
        set_error("%s: this context has no communicator (ndb_comm_init)", who);
        return NDB_ERR_INVALID_ARG;
    }
    return NDB_OK;
}

int64_t view_len(const ndb_view *v)
{
    int64_t n = 1;
    for (int i = 0; i < v->ndim; ++i) n *= v->dim[i];
    return n;
}

// a dense standard-layout view of `shape` over `ptr`
ndb_view dense_like(const ndb_view *shape, void *ptr)
{
    ndb_view v = *shape;
    v.ptr = ptr;
    int64_t st = 1;
    for (int i = v.ndim - 1; i >= 0; --i) {
        v.stride[i] = st;
        st *= v.dim[i];
    }
    return v;
}

}  // namespace

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_comm_unique_id(void *id_out /* 128 bytes */)
{
    clear_error();
    if (!id_out) return NDB_ERR_INVALID_ARG;
    if (!nccl().ok) {
        set_error("ndb_comm_unique_id: libnccl.so.2 could not be loaded (set NDB_NCCL_LIB)");
        return NDB_ERR_CUDA;
    }
    NcclUniqueId id;
    NDB_NCCL(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof(id));
    return NDB_OK;
}

int ndb_comm_init(ndb_ctx *ctx, int rank, int world, const void *id /* 128 bytes, from rank 0's ndb_comm_unique_id */)
{
    clear_error();
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) {
        set_error("ndb_comm_init: need a context, an id and 0 <= rank < world");
        return NDB_ERR_INVALID_ARG;
    }
    if (ctx->nccl_comm) {
        set_error("ndb_comm_init: this context already has a communicator");
        return NDB_ERR_INVALID_ARG;
    }
    if (!nccl().ok) {
        set_error("ndb_comm_init: libnccl.so.2 could not be loaded (set NDB_NCCL_LIB)");
        return NDB_ERR_CUDA;
    }
    DeviceGuard g(ctx->device);
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    nccl_comm_t comm = nullptr;
    NDB_NCCL(nccl().CommInitRank(&comm, world, uid, rank), "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return NDB_OK;
}

int ndb_comm_destroy(ndb_ctx *ctx)
{
    clear_error();
    if (!ctx || !ctx->nccl_comm) return NDB_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl().CommDestroy((nccl_comm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->comm_rank = -1;
    ctx->comm_world = 0;
    return NDB_OK;
}

int ndb_broadcast(ndb_ctx *ctx, const ndb_view *v, int src)
{
    clear_error();
    NDB_TRY(need_comm(ctx, "ndb_broadcast"));
    NDB_TRY(validate_view(v, "ndb_broadcast"));
    if (!ndb_view_is_contiguous(v)) {
        set_error("ndb_broadcast: the replicated operand must be one dense block (any axis order)");
        return NDB_ERR_INCOMPATIBLE_LAYOUT;
    }
    const int64_t n = view_len(v);
    if (n == 0) return NDB_OK;
    // lowest address of the block (negative strides put the logical first element elsewhere)
    int64_t low = 0;
    for (int i = 0; i < v->ndim; ++i)
        if (v->stride[i] < 0) low += (v->dim[i] - 1) * v->stride[i];
    void *base = (char *)v->ptr + low * (int64_t)dtype_size(v->dtype);
    DeviceGuard g(ctx->device);
    NDB_NCCL(nccl().Broadcast(base, base, (size_t)n, nccl_dtype(v->dtype), src, (nccl_comm_t)ctx->nccl_comm, ctx->stream), "ncclBroadcast");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NDB_OK;
}

int ndb_gemm_rowsharded(ndb_ctx *ctx, const void *alpha, const ndb_view *a_local, const ndb_view *b, const void *beta,
                        const ndb_view *c_local, int src)
{
    clear_error();
    if (ctx && ctx->comm_world <= 1 && !ctx->nccl_comm) return ndb_gemm(ctx, alpha, a_local, b, beta, c_local);  // one rank: nothing to exchange
    NDB_TRY(need_comm(ctx, "ndb_gemm_rowsharded"));
    if (ctx->comm_world > 1) NDB_TRY(ndb_broadcast(ctx, b, src));
    return ndb_gemm(ctx, alpha, a_local, b, beta, c_local);  // same stream: ordered after the broadcast
}

int ndb_reduce_all_sharded(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_host)
{
    clear_error();
    if (!ctx || !out_host || !in) {
        set_error("reduce_all_sharded: null context, input or output (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    if (redop < NDB_RED_SUM || redop > NDB_RED_MAX) {
        set_error("reduce_all_sharded: redop must be sum, product, min or max");
        return NDB_ERR_UNSUPPORTED;
    }
    DeviceGuard g(ctx->device);
    NDB_TRY(ensure_workspace(ctx, 4096));
    void *dev_scalar = (char *)ctx->workspace + ctx->workspace_bytes - 64;
    if (ctx->peer_world > 1 && ctx->peer_attached) {
        NDB_TRY(ndb_reduce_all_peer(ctx, redop, in, dev_scalar));
    } else {
        // an empty shard contributes the identity (min / max: the type's extreme) - it must not return before the exchange,
        // the other ranks would wait for it
        if (view_len(in) == 0 && ctx->nccl_comm && ctx->comm_world > 1) NDB_TRY(write_identity(ctx, in->dtype, redop, dev_scalar));
        else NDB_TRY(ndb_reduce_all_dev(ctx, redop, in, dev_scalar));
        if (ctx->nccl_comm && ctx->comm_world > 1) {
            static const int ops[4] = {NCCL_SUM, NCCL_PROD, NCCL_MIN, NCCL_MAX};
            NDB_NCCL(nccl().AllReduce(dev_scalar, dev_scalar, 1, nccl_dtype(in->dtype), ops[redop], (nccl_comm_t)ctx->nccl_comm, ctx->stream),
                     "ncclAllReduce");
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
    }
    NDB_CUDA(cudaMemcpyAsync(ctx->host_scalar, dev_scalar, dtype_size(in->dtype), cudaMemcpyDeviceToHost, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out_host, ctx->host_scalar, dtype_size(in->dtype));
    return NDB_OK;
}

int ndb_reduce_axis_sharded(ndb_ctx *ctx, int redop, int axis, const ndb_view *in, const ndb_view *out, double ddof, int64_t n_global)
{
    clear_error();
    if (!ctx || !in || !out) {
        set_error("reduce_axis_sharded: null context or view (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    if (redop < 0 || redop >= NDB_RED__COUNT) {
        set_error("reduce_axis_sharded: unknown redop %d", redop);
        return NDB_ERR_UNSUPPORTED;
    }
    const bool multi = (ctx->peer_world > 1 && ctx->peer_attached) || (ctx->nccl_comm && ctx->comm_world > 1);
    // a kept axis is reduced shard by shard: outputs are sharded like the input, nothing to exchange
    if (axis != 0 || !multi) return ndb_reduce_axis(ctx, redop, axis, in, out, ddof);
    if (n_global < in->dim[0]) {
        set_error("reduce_axis_sharded: n_global (%lld) is smaller than this shard's length", (long long)n_global);
        return NDB_ERR_INVALID_ARG;
    }
    // 1. the whole sharded reduction as one cooperative launch over the peer windows
    if (ctx->peer_world > 1 && ctx->peer_attached && n_global > 0) {
        const int st = ndb_reduce_axis_peer(ctx, redop, axis, in, out, ddof, n_global);
        if (st != NDB_ERR_UNSUPPORTED) return st;
        clear_error();
    }
    const int dt = in->dtype;
    const size_t esz = dtype_size(dt);
    const int64_t n_out = view_len(out);
    const int world = ctx->nccl_comm ? ctx->comm_world : ctx->peer_world;
    const bool welford = redop == NDB_RED_VAR || redop == NDB_RED_STD;
    if (welford && !dtype_is_float(dt)) {
        set_error("var_axis/std_axis need a float element type");
        return NDB_ERR_UNSUPPORTED;
    }
    if (welford && (ddof < 0.0 || ddof > (double)n_global)) {  // NaN passes, as in ndarray (impl_numeric.rs:408-413)
        set_error("`ddof` must not be less than zero or greater than the length of the axis");
        return NDB_ERR_INVALID_ARG;
    }
    if (redop == NDB_RED_MEAN && n_global == 0) {
        set_error("reduction of a zero-length axis has no value (ndarray returns None)");
        return NDB_ERR_EMPTY;
    }
    if (n_out == 0) return NDB_OK;
    DeviceGuard g(ctx->device);
    const bool out_dense = is_layout_c(out);
    if (!welford) {
        // 2. local partial (sum for mean), all-reduce over the ranks, mean divides by the GLOBAL length
        const int local_op = redop == NDB_RED_MEAN ? NDB_RED_SUM : redop;
        const bool minmax = local_op == NDB_RED_MIN || local_op == NDB_RED_MAX;
        if (minmax && n_global == 0) {  // the same on every rank: nobody enters the exchange
            set_error("reduction of a zero-length axis has no value (ndarray returns None)");
            return NDB_ERR_EMPTY;
        }
        void *tmp = nullptr;
        ndb_view part = *out;
        if (!out_dense) {
            NDB_CUDA(cudaMallocAsync(&tmp, (size_t)n_out * esz, ctx->stream));
            part = dense_like(out, tmp);
        }
        int st;
        if (minmax && in->dim[0] == 0) {
            // this rank holds no rows of a non-empty axis: it still takes part, contributing the type's extreme
            union { float f; double d; int32_t i; int64_t l; } ext;
            const bool mn = local_op == NDB_RED_MIN;
            switch (dt) {
            case NDB_F32: ext.f = mn ? INFINITY : -INFINITY; break;
            case NDB_F64: ext.d = mn ? INFINITY : -INFINITY; break;
            case NDB_I32: ext.i = mn ? INT32_MAX : INT32_MIN; break;
            default: ext.l = mn ? INT64_MAX : INT64_MIN; break;
            }
            st = ndb_zip(ctx, NDB_OP_FILL, &ext, 1, &part);
        } else {
            st = ndb_reduce_axis(ctx, local_op, 0, in, &part, 0.0);
        }
        if (st == NDB_OK) {
            if (ctx->nccl_comm) {
                static const int ops[4] = {NCCL_SUM, NCCL_PROD, NCCL_MIN, NCCL_MAX};
                const int rc = nccl().AllReduce(part.ptr, part.ptr, (size_t)n_out, nccl_dtype(dt), ops[local_op], (nccl_comm_t)ctx->nccl_comm, ctx->stream);
                if (rc != 0) st = nccl_fail(rc, "ncclAllReduce");
                g_launches.fetch_add(1, std::memory_order_relaxed);
            } else {
                st = ndb_peer_allreduce(ctx, local_op, dt, part.ptr, part.ptr, n_out);
            }
        }
        if (st == NDB_OK && redop == NDB_RED_MEAN) {
            union { float f; double d; int32_t i; int64_t l; } nT;
            switch (dt) {
            case NDB_F32: nT.f = (float)n_global; break;
            case NDB_F64: nT.d = (double)n_global; break;
            case NDB_I32: nT.i = (int32_t)n_global; break;
            default: nT.l = n_global; break;
            }
            ndb_view io[2] = {part, part};
            st = ndb_zip(ctx, NDB_OP_DIV_S, &nT, 2, io);
        }
        if (st == NDB_OK && !out_dense) {
            ndb_view io[2] = {*out, part};
            st = ndb_zip(ctx, NDB_OP_COPY, nullptr, 2, io);
        }
        if (tmp) cudaFreeAsync(tmp, ctx->stream);
        return st;
    }
    // 3. var / std: per-shard (n, mean, M2), all-gather, Chan merge in rank order
    char *buf = nullptr;
    const size_t vec = ((size_t)n_out * esz + 255) & ~(size_t)255;
    //   [mean | m2 | count | gathered means | gathered m2s | gathered counts | dense result]
    const size_t total = 2 * vec + 256 + 2 * (size_t)world * vec + 256 + (size_t)((world * 8 + 255) & ~255) + vec;
    NDB_CUDA(cudaMallocAsync((void **)&buf, total, ctx->stream));
    char *mean = buf, *m2 = buf + vec, *cnt = buf + 2 * vec, *gmean = cnt + 256, *gm2 = gmean + (size_t)world * vec;
    char *gcnt = gm2 + (size_t)world * vec, *res = gcnt + ((world * 8 + 255) & ~255);
    ndb_view vmean = dense_like(out, mean), vm2 = dense_like(out, m2);
    int st = ndb_welford_axis(ctx, 0, in, &vmean, &vm2);
    if (st == NDB_OK) {
        // this shard's row count, as a kernel parameter: the call returns without synchronising, and a copy staged through
        // ctx->host_scalar could still be pending when the next call overwrites that buffer
        const int64_t rows = in->dim[0];
        ndb_view vc;
        memset(&vc, 0, sizeof(vc));
        vc.ptr = cnt;
        vc.dtype = NDB_I64;
        vc.ndim = 1;
        vc.dim[0] = 1;
        vc.stride[0] = 1;
        st = ndb_zip(ctx, NDB_OP_FILL, &rows, 1, &vc);
    }
    if (st == NDB_OK) {
        if (ctx->nccl_comm) {
            nccl_comm_t comm = (nccl_comm_t)ctx->nccl_comm;
            // gathered vectors are padded to `vec` bytes per rank so that every rank's slice starts aligned
            int rc = nccl().AllGather(mean, gmean, vec / esz, nccl_dtype(dt), comm, ctx->stream);
            if (rc == 0) rc = nccl().AllGather(m2, gm2, vec / esz, nccl_dtype(dt), comm, ctx->stream);
            if (rc == 0) rc = nccl().AllGather(cnt, gcnt, 1, NCCL_INT64, comm, ctx->stream);
            if (rc != 0) st = nccl_fail(rc, "ncclAllGather");
            g_launches.fetch_add(3, std::memory_order_relaxed);
        } else {
            st = ndb_peer_allgather(ctx, dt, mean, gmean, (int64_t)(vec / esz));
            if (st == NDB_OK) st = ndb_peer_allgather(ctx, dt, m2, gm2, (int64_t)(vec / esz));
            if (st == NDB_OK) st = ndb_peer_allgather(ctx, NDB_I64, cnt, gcnt, 1);
        }
    }
    // gathered slices are `vec` bytes apart: the merge kernel takes that pitch in elements
    if (st == NDB_OK)
        st = welford_merge_gathered(ctx, dt, world, (const int64_t *)gcnt, gmean, gm2, n_out, (int64_t)(vec / esz), ddof, redop == NDB_RED_STD,
                                    out_dense ? out->ptr : res);
    if (st == NDB_OK && !out_dense) {
        ndb_view dense = dense_like(out, res);
        ndb_view io[2] = {*out, dense};
        st = ndb_zip(ctx, NDB_OP_COPY, nullptr, 2, io);
    }
    cudaFreeAsync(buf, ctx->stream);
    return st;
}

}  // extern "C"
