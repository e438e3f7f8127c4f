// ndb_zip: plan + launch; ndb_zip_host: the end-to-end form over HOST buffers.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "zip_kernels.cuh"

using namespace ndb;

namespace ndb {

static int zip_launch_plan(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    if (z.kernel == NDB_KERNEL_NOOP) return NDB_OK;  // Zip over an empty shape does nothing (tests/azip.rs:472)
    switch (z.dtype) {
    case NDB_F32: return zip_dispatch_f32(ctx, z, scalars);
    case NDB_F64: return zip_dispatch_f64(ctx, z, scalars);
    case NDB_I32: return zip_dispatch_i32(ctx, z, scalars);
    case NDB_I64: return zip_dispatch_i64(ctx, z, scalars);
    default: break;
    }
    set_error("Zip: unsupported dtype %d", z.dtype);
    return NDB_ERR_UNSUPPORTED;
}

static bool op_can_trap(int op, int dtype)
{
    if (dtype_is_float(dtype)) return false;
    return op == NDB_OP_DIV || op == NDB_OP_REM || op == NDB_OP_DIV_S || op == NDB_OP_REM_S || op == NDB_OP_RDIV_S;
}

// Integer / and % panic in Rust on a zero divisor or MIN / -1.  The kernels raise a device flag;
// read it back (these ops synchronise, the others stay asynchronous).
static int check_trap(ndb_ctx *ctx)
{
    NDB_CUDA(cudaMemcpyAsync(ctx->host_flag, ctx->dev_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NDB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*ctx->host_flag) {
        *ctx->host_flag = 0;
        NDB_CUDA(cudaMemsetAsync(ctx->dev_flag, 0, sizeof(int), ctx->stream));
        set_error("attempt to divide by zero (or to divide with overflow)");
        return NDB_ERR_DIV_BY_ZERO;
    }
    return NDB_OK;
}

}  // namespace ndb

extern "C" {

int ndb_zip(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts)
{
    if (!ctx) {
        set_error("ndb_zip: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    ZipPlan z;
    NDB_TRY(make_zip_plan(op, n_parts, parts, &z));
    DeviceGuard g(ctx->device);
    NDB_TRY(zip_launch_plan(ctx, z, scalars));
    if (z.kernel != NDB_KERNEL_NOOP && op_can_trap(op, z.dtype)) return check_trap(ctx);
    return NDB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Zip::all / any / fold (src/zip/mod.rs:622-696).  The closure's result is materialised into a dense scratch
// array by the Zip engine and reduced by the reduction engine (MIN for all, MAX for any, SUM for fold).
extern "C" int ndb_reduce_all(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_host);

namespace ndb {

static int zip_then_reduce(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int redop,
                           void *out_host, bool *empty)
{
    clear_error();
    *empty = false;
    if (!ctx) {
        set_error("Zip: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    if (n_inputs < 1 || n_inputs > NDB_MAX_PARTS - 1 || !inputs) {
        set_error("Zip: between 1 and %d input producers are supported", NDB_MAX_PARTS - 1);
        return NDB_ERR_UNSUPPORTED;
    }
    NDB_TRY(validate_view(&inputs[0], "Zip producer"));
    ndb_view parts[NDB_MAX_PARTS];
    ndb_view &out = parts[0];
    memset(&out, 0, sizeof(out));
    out.ndim = inputs[0].ndim;
    out.dtype = inputs[0].dtype;
    int64_t total = 1;
    for (int i = out.ndim - 1; i >= 0; --i) {
        out.dim[i] = inputs[0].dim[i];
        out.stride[i] = total;
        total *= out.dim[i];
    }
    for (int p = 0; p < n_inputs; ++p) parts[p + 1] = inputs[p];
    if (total == 0) {
        // still validate shapes / arity like a real traversal would
        ZipPlan z;
        out.ptr = inputs[0].ptr;
        NDB_TRY(make_zip_plan(op, n_inputs + 1, parts, &z));
        *empty = true;
        return NDB_OK;
    }
    DeviceGuard g(ctx->device);
    if (n_inputs == 2 && !scalars) {
        // both producers dense, identically laid out and aligned: one fused pass, no scratch
        ndb_view two[2] = {inputs[0], inputs[1]};
        ZipPlan z;
        if (make_zip_plan(NDB_OP_COPY, 2, two, &z) == NDB_OK && z.kernel == NDB_KERNEL_FLAT_VEC) {
            const int st = zip2_reduce_dense(ctx, op, redop, out.dtype, z.ptr[0], z.ptr[1], z.total, out_host);
            if (st != NDB_ERR_UNSUPPORTED) return st;
        }
        clear_error();
    }
    void *tmp = nullptr;
    NDB_CUDA(cudaMallocAsync(&tmp, (size_t)total * dtype_size(out.dtype), ctx->stream));
    out.ptr = tmp;
    int st = ndb_zip(ctx, op, scalars, n_inputs + 1, parts);
    if (st == NDB_OK) st = ndb_reduce_all(ctx, redop, &out, out_host);
    cudaFreeAsync(tmp, ctx->stream);
    return st;
}

static bool nonzero(int dtype, const void *v)
{
    switch (dtype) {
    case NDB_F32: return *(const float *)v != 0.f;
    case NDB_F64: return *(const double *)v != 0.0;
    case NDB_I32: return *(const int32_t *)v != 0;
    default: return *(const int64_t *)v != 0;
    }
}

static bool is_predicate(int op) { return op >= NDB_OP_EQ && op <= NDB_OP_GE; }

}  // namespace ndb

// ---- Zip::fold_while (src/zip/mod.rs:640-655) -----------------------------------------------------------------------
// FoldWhile::Done has a meaning only relative to a traversal ORDER, which a tree reduction does not have; the device form
// keeps the order explicit: the closure's values are materialised in Zip's traversal order (C, or F when the producers
// prefer it), the FIRST position where the stop predicate holds is found, and the fold runs over the prefix before it.
template <typename T>
__global__ void __launch_bounds__(256) first_nonzero_kernel(const T *flags, int64_t n, unsigned long long *first)
{
    pdl_enter();
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        if (flags[i] != T(0)) {
            atomicMin(first, (unsigned long long)i);
            break;  // this thread's later indices are larger
        }
    }
}

// acc = redop(acc, x) on the host, in the element type (wrapping integers, Rust's NaN-ignoring min / max)
template <typename T>
static void host_combine(int redop, T *acc, T x)
{
    using U = typename UnsignedOf<T>::type;
    switch (redop) {
    case NDB_RED_SUM:
        if constexpr (std::is_floating_point<T>::value) *acc = *acc + x;
        else *acc = (T)((U)*acc + (U)x);
        break;
    case NDB_RED_PROD:
        if constexpr (std::is_floating_point<T>::value) *acc = *acc * x;
        else *acc = (T)((U)*acc * (U)x);
        break;
    case NDB_RED_MIN: *acc = (x < *acc || *acc != *acc) ? x : *acc; break;
    default: *acc = (x > *acc || *acc != *acc) ? x : *acc; break;
    }
}

extern "C" {

int ndb_zip_all(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int *result)
{
    if (!result) return NDB_ERR_INVALID_ARG;
    if (!is_predicate(op)) {
        set_error("Zip::all takes a predicate op (NDB_OP_EQ..NDB_OP_GE)");
        return NDB_ERR_UNSUPPORTED;
    }
    unsigned char v[8] = {0};
    bool empty = false;
    NDB_TRY(zip_then_reduce(ctx, op, scalars, n_inputs, inputs, NDB_RED_MIN, v, &empty));
    *result = empty ? 1 : (nonzero(inputs[0].dtype, v) ? 1 : 0);  // vacuously true on an empty shape
    return NDB_OK;
}

int ndb_zip_any(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int *result)
{
    if (!result) return NDB_ERR_INVALID_ARG;
    if (!is_predicate(op)) {
        set_error("Zip::any takes a predicate op (NDB_OP_EQ..NDB_OP_GE)");
        return NDB_ERR_UNSUPPORTED;
    }
    unsigned char v[8] = {0};
    bool empty = false;
    NDB_TRY(zip_then_reduce(ctx, op, scalars, n_inputs, inputs, NDB_RED_MAX, v, &empty));
    *result = empty ? 0 : (nonzero(inputs[0].dtype, v) ? 1 : 0);
    return NDB_OK;
}

int ndb_zip_fold_sum(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, void *out_host)
{
    if (!out_host) return NDB_ERR_INVALID_ARG;
    bool empty = false;
    NDB_TRY(zip_then_reduce(ctx, op, scalars, n_inputs, inputs, NDB_RED_SUM, out_host, &empty));
    if (empty) memset(out_host, 0, dtype_size(inputs[0].dtype));
    return NDB_OK;
}

int ndb_zip_fold(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int redop, const void *init,
                 void *out_host)
{
    if (!out_host || redop < NDB_RED_SUM || redop > NDB_RED_MAX) {
        set_error("Zip::fold: the accumulator must be sum, product, min or max (an associative fold) and the result non-null");
        return NDB_ERR_INVALID_ARG;
    }
    if (!init && (redop == NDB_RED_MIN || redop == NDB_RED_MAX) && inputs && n_inputs >= 1) {
        int64_t total = 1;
        for (int i = 0; i < inputs[0].ndim; ++i) total *= inputs[0].dim[i];
        if (total == 0) {
            set_error("Zip::fold(min/max) over an empty shape needs an initial value");
            return NDB_ERR_EMPTY;
        }
    }
    bool empty = false;
    unsigned char v[8] = {0};
    NDB_TRY(zip_then_reduce(ctx, op, scalars, n_inputs, inputs, redop, v, &empty));
    const int dt = inputs[0].dtype;
    const size_t esz = dtype_size(dt);
    if (empty) {  // fold over nothing returns the initial accumulator (src/zip/mod.rs:622); without one, the identity
        if (init) memcpy(out_host, init, esz);
        else {
            memset(out_host, 0, esz);
            if (redop == NDB_RED_PROD) {
                switch (dt) {
                case NDB_F32: *(float *)out_host = 1.f; break;
                case NDB_F64: *(double *)out_host = 1.0; break;
                case NDB_I32: *(int32_t *)out_host = 1; break;
                default: *(int64_t *)out_host = 1; break;
                }
            }
        }
        return NDB_OK;
    }
    if (!init) {
        memcpy(out_host, v, esz);
        return NDB_OK;
    }
    memcpy(out_host, init, esz);
    switch (dt) {
    case NDB_F32: host_combine<float>(redop, (float *)out_host, *(const float *)v); break;
    case NDB_F64: host_combine<double>(redop, (double *)out_host, *(const double *)v); break;
    case NDB_I32: host_combine<int32_t>(redop, (int32_t *)out_host, *(const int32_t *)v); break;
    default: host_combine<int64_t>(redop, (int64_t *)out_host, *(const int64_t *)v); break;
    }
    return NDB_OK;
}

int ndb_zip_fold_while(ndb_ctx *ctx, int op, const void *scalars, int n_inputs, const ndb_view *inputs, int redop, const void *init,
                       int stop_op, const void *stop_value, void *out_host, int *done)
{
    clear_error();
    if (!ctx || !out_host || !done || !inputs || !stop_value || redop < NDB_RED_SUM || redop > NDB_RED_MAX || !is_predicate(stop_op)) {
        set_error("Zip::fold_while: need an associative accumulator (sum/prod/min/max), a stop predicate (NDB_OP_EQ..GE) with its value, and result pointers");
        return NDB_ERR_INVALID_ARG;
    }
    if (n_inputs < 1 || n_inputs > NDB_MAX_PARTS - 1) {
        set_error("Zip: between 1 and %d input producers are supported", NDB_MAX_PARTS - 1);
        return NDB_ERR_UNSUPPORTED;
    }
    *done = 0;
    NDB_TRY(validate_view(&inputs[0], "Zip producer"));
    const int dt = inputs[0].dtype;
    const size_t esz = dtype_size(dt);
    // scratch in traversal order: F-order iff the producers prefer it (src/zip/mod.rs:278-281)
    const bool f_order = ndb_zip_prefer_f(n_inputs, inputs) != 0;
    ndb_view parts[NDB_MAX_PARTS];
    ndb_view &vals = parts[0];
    memset(&vals, 0, sizeof(vals));
    vals.ndim = inputs[0].ndim;
    vals.dtype = dt;
    int64_t total = 1;
    for (int k = 0; k < vals.ndim; ++k) {
        const int i = f_order ? k : vals.ndim - 1 - k;
        vals.dim[i] = inputs[0].dim[i];
        vals.stride[i] = total;
        total *= vals.dim[i];
    }
    for (int p = 0; p < n_inputs; ++p) parts[p + 1] = inputs[p];
    if (total == 0) {  // nothing to visit: the initial accumulator, not Done
        vals.ptr = inputs[0].ptr;
        ZipPlan z;
        NDB_TRY(make_zip_plan(op, n_inputs + 1, parts, &z));
        return ndb_zip_fold(ctx, op, scalars, n_inputs, inputs, redop, init, out_host);
    }
    DeviceGuard g(ctx->device);
    char *buf = nullptr;  // [values | flags | stop value | first index]
    const size_t vec = ((size_t)total * esz + 255) & ~(size_t)255;
    NDB_CUDA(cudaMallocAsync((void **)&buf, 2 * vec + 512, ctx->stream));
    vals.ptr = buf;
    int st = ndb_zip(ctx, op, scalars, n_inputs + 1, parts);
    unsigned long long first = (unsigned long long)total;
    if (st == NDB_OK) {
        ndb_view flat = vals, flags, sv;
        flat.ndim = 1;
        flat.dim[0] = total;
        flat.stride[0] = 1;
        flags = flat;
        flags.ptr = buf + vec;
        sv = flat;
        sv.ptr = buf + 2 * vec;
        sv.stride[0] = 0;  // the stop value, broadcast
        memcpy(ctx->host_scalar, stop_value, esz);
        *(unsigned long long *)((char *)ctx->host_scalar + 8) = first;
        if (cudaMemcpyAsync(buf + 2 * vec, ctx->host_scalar, 16, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) st = NDB_ERR_CUDA;
        ndb_view cmp[3] = {flags, flat, sv};
        if (st == NDB_OK) st = ndb_zip(ctx, stop_op, nullptr, 3, cmp);
        if (st == NDB_OK) {
            unsigned long long *dev_first = (unsigned long long *)(buf + 2 * vec + 8);
            const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 8, (total + 255) / 256));
            switch (dt) {
            case NDB_F32: launch_pdl(first_nonzero_kernel<float>, blocks, 256, 0, ctx->stream, (const float *)flags.ptr, total, dev_first); break;
            case NDB_F64: launch_pdl(first_nonzero_kernel<double>, blocks, 256, 0, ctx->stream, (const double *)flags.ptr, total, dev_first); break;
            case NDB_I32: launch_pdl(first_nonzero_kernel<int32_t>, blocks, 256, 0, ctx->stream, (const int32_t *)flags.ptr, total, dev_first); break;
            default: launch_pdl(first_nonzero_kernel<int64_t>, blocks, 256, 0, ctx->stream, (const int64_t *)flags.ptr, total, dev_first); break;
            }
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaMemcpyAsync(ctx->host_scalar, dev_first, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                st = NDB_ERR_CUDA;
            else first = *(unsigned long long *)ctx->host_scalar;
        }
        if (st == NDB_OK) {
            *done = first < (unsigned long long)total ? 1 : 0;
            flat.dim[0] = (int64_t)first;  // the prefix before the stopping element
            ndb_view one[1] = {flat};
            st = ndb_zip_fold(ctx, NDB_OP_COPY, nullptr, 1, one, redop, init, out_host);
            if (st == NDB_ERR_EMPTY) {  // min / max over an empty prefix without an initial value
                set_error("Zip::fold_while(min/max) stopped before the first element and no initial value was given");
            }
        }
    }
    cudaFreeAsync(buf, ctx->stream);
    return st;
}


}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Host-buffer path.  Every producer lives in host memory; each one's address span is mirrored in a
// device staging buffer.  When the outermost axis separates a producer into disjoint slabs, slabs are
// streamed: H2D of slab c+1, the kernel on slab c and D2H of slab c-1 overlap on three streams, so
// the call costs about max(bytes in, bytes out) / PCIe bandwidth instead of their sum.
namespace {

struct Span {
    int64_t lo, hi;  // element offsets (inclusive) relative to view.ptr; empty if hi < lo
};

Span view_span(const ndb_view &v, int from_axis = 0)
{
    Span s{0, 0};
    for (int k = from_axis; k < v.ndim; ++k) {
        if (v.dim[k] == 0) return Span{0, -1};
        const int64_t e = (v.dim[k] - 1) * v.stride[k];
        if (e < 0) s.lo += e;
        else s.hi += e;
    }
    return s;
}

// plain malloc / Vec memory (not cudaHostAlloc'ed, not registered)?
bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

bool slabs_disjoint(const ndb_view &v)
{
    if (v.ndim == 0 || v.dim[0] <= 1) return false;
    const Span inner = view_span(v, 1);
    return llabs(v.stride[0]) >= (inner.hi - inner.lo + 1);
}

}  // namespace

extern "C" int ndb_reduce_all_dev(ndb_ctx *ctx, int redop, const ndb_view *in, void *out_dev);

// redop < 0: plain Zip.  Otherwise every slab's output is also reduced on the device while it is still resident (one
// scalar per slab, in slab order), and the reduction of those scalars is written to `fold_out` (host, element type):
// Zip::fold over the values the Zip has just produced, with no second pass over host memory.
static int zip_host_impl(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts, int redop, void *fold_out)
{
    if (!ctx) {
        set_error("ndb_zip_host: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    ZipPlan probe;
    NDB_TRY(make_zip_plan(op, n_parts, parts, &probe));  // validates shapes, arity, aliasing
    const int64_t esz = (int64_t)dtype_size(parts[0].dtype);
    if (probe.kernel == NDB_KERNEL_NOOP) {
        if (redop >= 0) {  // fold over nothing: sum -> 0, product -> 1; min / max have no value
            if (redop != NDB_RED_SUM && redop != NDB_RED_PROD) {
                set_error("Zip::fold(min/max) over an empty shape has no value");
                return NDB_ERR_EMPTY;
            }
            ndb_view none = parts[0];
            return ndb_reduce_all(ctx, redop, &none, fold_out);
        }
        return NDB_OK;
    }
    DeviceGuard g(ctx->device);

    // inputs that alias the output (a += b passes `a` twice) ride on the output's buffer
    bool alias_out[NDB_MAX_PARTS] = {false};
    for (int p = 1; p < n_parts; ++p)
        alias_out[p] = parts[p].ptr == parts[0].ptr && memcmp(parts[p].dim, parts[0].dim, sizeof(parts[0].dim)) == 0 &&
                       memcmp(parts[p].stride, parts[0].stride, sizeof(parts[0].stride)) == 0;
    // mirror every other producer's span on the device
    ndb_view dev[NDB_MAX_PARTS];
    Span span[NDB_MAX_PARTS];
    for (int p = 0; p < n_parts; ++p) {
        dev[p] = parts[p];
        if (alias_out[p]) {
            dev[p].ptr = dev[0].ptr;
            continue;
        }
        span[p] = view_span(parts[p]);
        const size_t bytes = (size_t)((span[p].hi - span[p].lo + 1) * esz);
        NDB_TRY(ensure_stage(ctx, p, bytes));
        dev[p].ptr = (char *)ctx->stage[p] - span[p].lo * esz;
    }
    const bool out_dense = ndb_view_is_contiguous(&parts[0]) != 0;
    bool out_is_read = !out_dense;
    for (int p = 1; p < n_parts; ++p) out_is_read |= alias_out[p];

    // can the outermost axis be streamed?
    bool stream_out = slabs_disjoint(parts[0]);
    int64_t n0 = parts[0].ndim > 0 ? parts[0].dim[0] : 1;
    int64_t total_bytes = probe.total * esz;
    int chunks = 1;
    if (stream_out) {
        chunks = (int)std::min<int64_t>(std::min<int64_t>(n0, 32), std::max<int64_t>(1, total_bytes / (16 << 20)));
    }
    cudaStream_t s_in = ctx->copy_in, s_k = ctx->stream, s_out = ctx->copy_out;
    std::vector<cudaEvent_t> ev_in(chunks), ev_k(chunks);
    cudaEvent_t ev_start, ev_whole;
    NDB_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    NDB_CUDA(cudaEventCreateWithFlags(&ev_whole, cudaEventDisableTiming));
    for (int c = 0; c < chunks; ++c) {
        NDB_CUDA(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        NDB_CUDA(cudaEventCreateWithFlags(&ev_k[c], cudaEventDisableTiming));
    }
    // copies must not start before earlier work on the compute stream that may still use the staging buffers
    NDB_CUDA(cudaEventRecord(ev_start, s_k));
    NDB_CUDA(cudaStreamWaitEvent(s_in, ev_start, 0));
    NDB_CUDA(cudaStreamWaitEvent(s_out, ev_start, 0));

    // Move one view's elements: a pitched 2-D copy when axis 0 separates it into disjoint slabs (only the
    // slabs travel, e.g. every other plane of s![..;-2, .., ..]), else its whole address span.
    auto copy_view = [&](const ndb_view &hv, const ndb_view &dv, cudaMemcpyKind kind, cudaStream_t st) -> int {
        const Span sp = view_span(hv);
        if (sp.hi < sp.lo) return NDB_OK;
        const bool h2d = kind == cudaMemcpyHostToDevice;
        if (hv.ndim >= 2 && hv.dim[0] > 1) {
            const Span inner = view_span(hv, 1);
            const int64_t width = inner.hi - inner.lo + 1;
            const int64_t s0 = llabs(hv.stride[0]);
            if (s0 > width && s0 * esz < ((int64_t)1 << 31)) {
                const int64_t first = (hv.stride[0] < 0 ? (hv.dim[0] - 1) * hv.stride[0] : 0) + inner.lo;
                char *h = (char *)hv.ptr + first * esz;
                char *d = (char *)dv.ptr + first * esz;
                NDB_CUDA(cudaMemcpy2DAsync(h2d ? d : h, (size_t)(s0 * esz), h2d ? h : d, (size_t)(s0 * esz),
                                           (size_t)(width * esz), (size_t)hv.dim[0], kind, st));
                return NDB_OK;
            }
        }
        const size_t bytes = (size_t)((sp.hi - sp.lo + 1) * esz);
        char *h = (char *)hv.ptr + sp.lo * esz;
        char *d = (char *)dv.ptr + sp.lo * esz;
        if (bytes >= (size_t(4) << 20) && is_pageable(h)) {
            // ndarray's own storage (a Vec) is pageable: a plain cudaMemcpyAsync from it runs at ~12 GB/s and blocks; the
            // threaded pinned-bounce pipeline of the CBLAS boundary moves it at the host's memcpy rate instead
            return h2d ? staged_h2d(ctx, d, h, bytes, st) : staged_d2h(ctx, h, d, bytes, st);
        }
        NDB_CUDA(cudaMemcpyAsync(h2d ? d : h, h2d ? h : d, bytes, kind, st));
        return NDB_OK;
    };

    // producers whose slabs interleave along axis 0 go over whole, up front
    bool streamed[NDB_MAX_PARTS] = {false};
    for (int p = 1; p < n_parts; ++p) {
        if (alias_out[p]) continue;
        streamed[p] = chunks > 1 && slabs_disjoint(parts[p]);
        if (!streamed[p]) NDB_TRY(copy_view(parts[p], dev[p], cudaMemcpyHostToDevice, s_in));
    }
    if (chunks == 1 && out_is_read) NDB_TRY(copy_view(parts[0], dev[0], cudaMemcpyHostToDevice, s_in));
    NDB_CUDA(cudaEventRecord(ev_whole, s_in));
    NDB_CUDA(cudaStreamWaitEvent(s_k, ev_whole, 0));

    char *fold_part = nullptr;  // one scalar per slab (8-byte slots)
    if (redop >= 0) NDB_CUDA(cudaMallocAsync((void **)&fold_part, (size_t)chunks * 8, s_k));
    int status = NDB_OK;
    for (int c = 0; c < chunks && status == NDB_OK; ++c) {
        const int64_t lo = n0 * c / chunks, hi = n0 * (c + 1) / chunks;
        ndb_view hv[NDB_MAX_PARTS], dv[NDB_MAX_PARTS];
        for (int p = 0; p < n_parts; ++p) {
            hv[p] = parts[p];
            dv[p] = dev[p];
            if (chunks > 1) {
                hv[p].dim[0] = dv[p].dim[0] = hi - lo;
                hv[p].ptr = (char *)parts[p].ptr + lo * parts[p].stride[0] * esz;
                dv[p].ptr = (char *)dev[p].ptr + lo * dev[p].stride[0] * esz;
            }
        }
        if (chunks > 1) {
            for (int p = 1; p < n_parts && status == NDB_OK; ++p)
                if (streamed[p]) status = copy_view(hv[p], dv[p], cudaMemcpyHostToDevice, s_in);
            if (status == NDB_OK && out_is_read)
                status = copy_view(hv[0], dv[0], cudaMemcpyHostToDevice, s_in);
            if (status != NDB_OK) break;
            NDB_CUDA(cudaEventRecord(ev_in[c], s_in));
            NDB_CUDA(cudaStreamWaitEvent(s_k, ev_in[c], 0));
        }
        ZipPlan z;
        status = make_zip_plan(op, n_parts, dv, &z);
        if (status == NDB_OK) status = zip_launch_plan(ctx, z, scalars);
        if (status != NDB_OK) break;
        if (redop >= 0) {
            status = ndb_reduce_all_dev(ctx, redop, &dv[0], fold_part + (size_t)c * esz);
            if (status != NDB_OK) break;
        }
        NDB_CUDA(cudaEventRecord(ev_k[c], s_k));
        NDB_CUDA(cudaStreamWaitEvent(s_out, ev_k[c], 0));
        status = copy_view(hv[0], dv[0], cudaMemcpyDeviceToHost, s_out);
    }
    cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(s_k), e3 = cudaStreamSynchronize(s_out);
    cudaEventDestroy(ev_start);
    cudaEventDestroy(ev_whole);
    for (int c = 0; c < chunks; ++c) {
        cudaEventDestroy(ev_in[c]);
        cudaEventDestroy(ev_k[c]);
    }
    if (status == NDB_OK && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && redop >= 0) {
        ndb_view pv;
        memset(&pv, 0, sizeof(pv));
        pv.ptr = fold_part;
        pv.ndim = 1;
        pv.dtype = parts[0].dtype;
        pv.dim[0] = chunks;
        pv.stride[0] = 1;
        status = ndb_reduce_all(ctx, redop, &pv, fold_out);  // synchronises
    }
    if (fold_part) cudaFreeAsync(fold_part, s_k);
    if (status != NDB_OK) return status;
    NDB_CUDA(e1);
    NDB_CUDA(e2);
    NDB_CUDA(e3);
    if (op_can_trap(op, parts[0].dtype)) return check_trap(ctx);
    return NDB_OK;
}

extern "C" {

int ndb_zip_host(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts)
{
    return zip_host_impl(ctx, op, scalars, n_parts, parts, -1, nullptr);
}

int ndb_zip_host_fold(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts, int redop, void *out_host)
{
    if (!out_host || redop < NDB_RED_SUM || redop > NDB_RED_MAX) {
        set_error("ndb_zip_host_fold: redop must be sum, product, min or max and the result pointer non-null");
        return NDB_ERR_INVALID_ARG;
    }
    return zip_host_impl(ctx, op, scalars, n_parts, parts, redop, out_host);
}

}  // extern "C"
