// Index-addressable synthetic data ("generated on device per shard from the counter RNG, so any sharding
// yields identical data").  Harness plumbing like the event helpers, not part of ndarray's path: element i of a logical
// 1-D stream is a pure function of (seed, i), so a rank fills rows [lo, hi) of a global array by passing the flat index
// of its first element, and the CPU oracle (oracle/ndarray_oracle.cpp orc_counter_fill) reproduces any block bit for bit.
//
//   h(seed, i) = splitmix64 output number i of a generator seeded with `seed`
//   f32: top 24 bits -> k * 2^-23 - 1      in [-1, 1), exactly representable, mean -2^-24
//   f64: top 53 bits -> k * 2^-52 - 1      in [-1, 1)
//   i32: top 32 bits, i64: all 64 bits (full range, so wrapping sums are exercised)
#include "zip_kernels.cuh"

namespace ndb {

namespace {

__host__ __device__ __forceinline__ uint64_t counter_hash(uint64_t seed, uint64_t i)
{
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__device__ __forceinline__ T counter_value(uint64_t seed, uint64_t i)
{
    const uint64_t h = counter_hash(seed, i);
    if constexpr (std::is_same<T, float>::value) return (float)(uint32_t)(h >> 40) * 0x1p-23f - 1.0f;
    else if constexpr (std::is_same<T, double>::value) return (double)(h >> 11) * 0x1p-52 - 1.0;
    else if constexpr (std::is_same<T, int32_t>::value) return (int32_t)(uint32_t)(h >> 32);
    else return (int64_t)h;
}

// dense run, 128-bit stores; the base pointer is 16-byte aligned (checked on the host), the tail is scalar
template <typename T>
__global__ void __launch_bounds__(256) counter_fill(T *dst, int64_t n, uint64_t seed, uint64_t first)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    const int64_t nv = n / N;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nv; i += stride) {
        V v;
#pragma unroll
        for (int e = 0; e < N; ++e) v.v[e] = counter_value<T>(seed, first + (uint64_t)(i * N + e));
        reinterpret_cast<V *>(dst)[i] = v;
    }
    if (blockIdx.x == 0) {
        const int64_t t = nv * N + threadIdx.x;
        if (t < n) dst[t] = counter_value<T>(seed, first + (uint64_t)t);
    }
}

}  // namespace

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_fill_counter_rng(ndb_ctx *ctx, const ndb_view *out, uint64_t seed, uint64_t first_index)
{
    clear_error();
    if (!ctx) {
        set_error("fill_counter_rng: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    NDB_TRY(validate_view(out, "fill_counter_rng"));
    if (!is_layout_c(out)) {
        set_error("fill_counter_rng: the view must be standard-layout (the stream index is the C-order flat index)");
        return NDB_ERR_INCOMPATIBLE_LAYOUT;
    }
    int64_t n = 1;
    for (int i = 0; i < out->ndim; ++i) n *= out->dim[i];
    if (n == 0) return NDB_OK;
    if ((uintptr_t)out->ptr % 16 != 0) {
        set_error("fill_counter_rng: the base pointer must be 16-byte aligned");
        return NDB_ERR_INCOMPATIBLE_LAYOUT;
    }
    DeviceGuard g(ctx->device);
    const int64_t nv = n / (16 / (int64_t)dtype_size(out->dtype));
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 16, (nv + 255) / 256));
    switch (out->dtype) {
    case NDB_F32: launch_pdl(counter_fill<float>, blocks, 256, 0, ctx->stream, (float *)out->ptr, n, seed, first_index); break;
    case NDB_F64: launch_pdl(counter_fill<double>, blocks, 256, 0, ctx->stream, (double *)out->ptr, n, seed, first_index); break;
    case NDB_I32: launch_pdl(counter_fill<int32_t>, blocks, 256, 0, ctx->stream, (int32_t *)out->ptr, n, seed, first_index); break;
    case NDB_I64: launch_pdl(counter_fill<int64_t>, blocks, 256, 0, ctx->stream, (int64_t *)out->ptr, n, seed, first_index); break;
    default: return NDB_ERR_UNSUPPORTED;
    }
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}
