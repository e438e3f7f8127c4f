// ndb_cast: out[i] = in[i] as OutType over two views of equal shape and any signed strides.
// This is `a.mapv(|x| x as B)` (src/impl_methods.rs:2990 with an `as` closure): the one Zip form whose producers
// have different element types.  Rust's `as` semantics: float -> int truncates toward zero and SATURATES, NaN -> 0
// (which is exactly what PTX cvt.rzi does); int -> float rounds to nearest; int -> int wraps; f64 -> f32 rounds to
// nearest.  Geometry comes from the Zip planner (collapsed [batch][row][inner]); the kernel is the scalar rows form.
#include <type_traits>

#include "zip_plan.h"

namespace ndb {

namespace {

struct CastArgs {
    char *out;
    const char *in;
    int64_t bdim[NDB_MAX_NDIM - 2], bso[NDB_MAX_NDIM - 2], bsi[NDB_MAX_NDIM - 2];
    int n_batch;
    int64_t n_rows, rso, rsi;
    int64_t inner_len, iso, isi;
    int64_t blocks_inner, blocks_rows, total_blocks;
};

template <typename TO, typename TI>
__device__ __forceinline__ TO cast_one(TI x)
{
    if constexpr (std::is_floating_point<TI>::value && std::is_integral<TO>::value) {
        if (x != x) return TO(0);  // NaN -> 0 (cvt.rzi.s64 would give i64::MIN)
        if constexpr (sizeof(TO) == 4) return (TO)__double2int_rz((double)x);  // saturating
        else return (TO)__double2ll_rz((double)x);
    } else {
        return (TO)x;
    }
}

template <typename TO, typename TI>
__global__ void __launch_bounds__(256) cast_rows(CastArgs a)
{
    pdl_enter();
    const int TX = blockDim.x, TY = blockDim.y;
    for (int64_t vb = blockIdx.x; vb < a.total_blocks; vb += gridDim.x) {
        int64_t t = vb;
        const int64_t bin = t % a.blocks_inner;
        t /= a.blocks_inner;
        const int64_t brow = t % a.blocks_rows;
        t /= a.blocks_rows;
        int64_t oo = 0, oi = 0;
        for (int k = a.n_batch - 1; k >= 0; --k) {
            const int64_t q = t / a.bdim[k];
            const int64_t idx = t - q * a.bdim[k];
            t = q;
            oo += idx * a.bso[k];
            oi += idx * a.bsi[k];
        }
        const int64_t row = brow * TY + threadIdx.y;
        if (row >= a.n_rows) continue;
#pragma unroll 4
        for (int it = 0; it < 4; ++it) {
            const int64_t i = (bin * 4 + it) * TX + threadIdx.x;
            if (i < a.inner_len) {
                const TI x = reinterpret_cast<const TI *>(a.in)[oi + row * a.rsi + i * a.isi];
                reinterpret_cast<TO *>(a.out)[oo + row * a.rso + i * a.iso] = cast_one<TO, TI>(x);
            }
        }
    }
}

template <typename TO, typename TI>
int launch_cast(ndb_ctx *ctx, CastArgs &a)
{
    int tx = 256;
    while (tx > 1 && tx / 2 >= a.inner_len) tx /= 2;
    const int ty = 256 / tx;
    a.blocks_inner = (a.inner_len + (int64_t)tx * 4 - 1) / ((int64_t)tx * 4);
    a.blocks_rows = (a.n_rows + ty - 1) / ty;
    int64_t batches = 1;
    for (int k = 0; k < a.n_batch; ++k) batches *= a.bdim[k];
    a.total_blocks = a.blocks_inner * a.blocks_rows * batches;
    const unsigned blocks = (unsigned)std::min<int64_t>(a.total_blocks, (int64_t)1 << 30);
    launch_pdl(cast_rows<TO, TI>, blocks, dim3(tx, ty, 1), 0, ctx->stream, a);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

template <typename TO>
int dispatch_in(ndb_ctx *ctx, int in_dtype, CastArgs &a)
{
    switch (in_dtype) {
    case NDB_F32: return launch_cast<TO, float>(ctx, a);
    case NDB_F64: return launch_cast<TO, double>(ctx, a);
    case NDB_I32: return launch_cast<TO, int32_t>(ctx, a);
    case NDB_I64: return launch_cast<TO, int64_t>(ctx, a);
    }
    return NDB_ERR_UNSUPPORTED;
}

}  // namespace

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_cast(ndb_ctx *ctx, const ndb_view *out, const ndb_view *in)
{
    clear_error();
    if (!ctx) {
        set_error("cast: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    NDB_TRY(validate_view(out, "cast output"));
    NDB_TRY(validate_view(in, "cast input"));
    // plan the geometry as a same-typed copy (shape check, aliasing check, collapse), then fix up the pointers
    ndb_view parts[2] = {*out, *in};
    parts[1].dtype = out->dtype;
    ZipPlan z;
    NDB_TRY(make_zip_plan(NDB_OP_COPY, 2, parts, &z));
    if (z.kernel == NDB_KERNEL_NOOP) return NDB_OK;
    CastArgs a;
    memset(&a, 0, sizeof(a));
    a.out = (char *)out->ptr + z.offset[0] * (int64_t)dtype_size(out->dtype);
    a.in = (const char *)in->ptr + z.offset[1] * (int64_t)dtype_size(in->dtype);
    const int inner = z.ndim - 1;
    int row_axis = -1;
    for (int k = 0; k < inner; ++k)
        if (row_axis < 0 || z.dim[k] > z.dim[row_axis]) row_axis = k;
    a.inner_len = z.dim[inner];
    a.iso = z.stride[0][inner];
    a.isi = z.stride[1][inner];
    a.n_rows = row_axis >= 0 ? z.dim[row_axis] : 1;
    a.rso = row_axis >= 0 ? z.stride[0][row_axis] : 0;
    a.rsi = row_axis >= 0 ? z.stride[1][row_axis] : 0;
    for (int k = 0; k < inner; ++k) {
        if (k == row_axis) continue;
        a.bdim[a.n_batch] = z.dim[k];
        a.bso[a.n_batch] = z.stride[0][k];
        a.bsi[a.n_batch] = z.stride[1][k];
        ++a.n_batch;
    }
    DeviceGuard g(ctx->device);
    switch (out->dtype) {
    case NDB_F32: return dispatch_in<float>(ctx, in->dtype, a);
    case NDB_F64: return dispatch_in<double>(ctx, in->dtype, a);
    case NDB_I32: return dispatch_in<int32_t>(ctx, in->dtype, a);
    case NDB_I64: return dispatch_in<int64_t>(ctx, in->dtype, a);
    }
    return NDB_ERR_UNSUPPORTED;
}
