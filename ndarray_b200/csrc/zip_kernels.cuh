// Elementwise Zip kernels for sm_100a.  Three kernel families, picked by the planner (zip_plan.cu):
//
//   zip_flat_vec   every producer is one dense run: 128-bit ld/st, UNROLL independent vectors per
//                  thread per producer in flight (HBM-latency x bandwidth needs ~40 KB in flight per SM).
//   zip_rows       collapsed N-d views with arbitrary SIGNED strides.  Warps walk the output's
//                  fastest axis; the outer multi-index is decomposed once per thread-row, not per
//                  element.  VEC=true is the 128-bit form for aligned unit/zero inner strides
//                  (covers outer-axis slices such as s![..;-2, .., ..] and row broadcasts).
//   zip_tile       a producer is dense along ANOTHER axis than the output (x.t()): it is read in
//                  coalesced runs along its own dense axis into padded shared memory and consumed
//                  transposed, so both sides move full 32-byte sectors.
//
// Arithmetic contract (bit-exact vs ndarray on the CPU): this file is compiled with -fmad=false, so
// b*c+a rounds twice like Rust; integer +,-,* wrap (done in unsigned); integer / and % raise the
// device flag where Rust panics.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "common.h"
#include "zip_plan.h"

namespace ndb {

template <typename T>
struct UnsignedOf { using type = T; };
template <>
struct UnsignedOf<int32_t> { using type = uint32_t; };
template <>
struct UnsignedOf<int64_t> { using type = uint64_t; };

template <typename T>
struct alignas(16) Vec128 {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};

template <typename T>
__device__ __forceinline__ T w_add(T a, T b)
{
    using U = typename UnsignedOf<T>::type;
    return (T)((U)a + (U)b);
}
template <typename T>
__device__ __forceinline__ T w_sub(T a, T b)
{
    using U = typename UnsignedOf<T>::type;
    return (T)((U)a - (U)b);
}
template <typename T>
__device__ __forceinline__ T w_mul(T a, T b)
{
    using U = typename UnsignedOf<T>::type;
    return (T)((U)a * (U)b);
}
template <typename T>
__device__ __forceinline__ T w_neg(T a)
{
    if constexpr (std::is_floating_point<T>::value) return -a;
    else {
        using U = typename UnsignedOf<T>::type;
        return (T)((U)0 - (U)a);
    }
}
template <typename T>
__device__ __forceinline__ T w_div(T a, T b, int *flag)
{
    if constexpr (std::is_floating_point<T>::value) return a / b;
    else {
        // Rust: "attempt to divide by zero" / "attempt to divide with overflow"
        if (b == 0 || (b == (T)-1 && a == (T)((typename UnsignedOf<T>::type)1 << (sizeof(T) * 8 - 1)))) {
            *flag = 1;
            return 0;
        }
        return a / b;
    }
}
template <typename T>
__device__ __forceinline__ T w_rem(T a, T b, int *flag)
{
    if constexpr (std::is_same<T, float>::value) return fmodf(a, b);
    else if constexpr (std::is_same<T, double>::value) return fmod(a, b);
    else {
        if (b == 0 || (b == (T)-1 && a == (T)((typename UnsignedOf<T>::type)1 << (sizeof(T) * 8 - 1)))) {
            *flag = 1;
            return 0;
        }
        return a % b;
    }
}
template <typename T>
__device__ __forceinline__ T w_abs(T a)
{
    if constexpr (std::is_same<T, float>::value) return fabsf(a);
    else if constexpr (std::is_same<T, double>::value) return fabs(a);
    else return a < 0 ? w_neg(a) : a;
}
// min/max: f32::min / f32::max ignore a NaN operand; the sign of a zero result is not pinned by ndarray,
// so both the kernels and the oracle use the same explicit rule (first operand wins ties).
template <typename T>
__device__ __forceinline__ T w_min(T a, T b)
{
    if constexpr (std::is_floating_point<T>::value) return (b < a || a != a) ? b : a;
    else return b < a ? b : a;
}
template <typename T>
__device__ __forceinline__ T w_max(T a, T b)
{
    if constexpr (std::is_floating_point<T>::value) return (b > a || a != a) ? b : a;
    else return b > a ? b : a;
}

// The float maps of src/numeric/impl_float_maths.rs, selected at run time (uniform across the launch).
template <typename T>
__device__ __forceinline__ T math_fn(int fn, T x, T p)
{
    if constexpr (std::is_floating_point<T>::value) {
        constexpr bool F = std::is_same<T, float>::value;
        switch (fn) {
        case NDB_FN_FLOOR: return F ? (T)floorf((float)x) : (T)floor((double)x);
        case NDB_FN_CEIL: return F ? (T)ceilf((float)x) : (T)ceil((double)x);
        case NDB_FN_ROUND: return F ? (T)roundf((float)x) : (T)round((double)x);  // half away from zero, like Rust
        case NDB_FN_TRUNC: return F ? (T)truncf((float)x) : (T)trunc((double)x);
        case NDB_FN_FRACT: return x - (F ? (T)truncf((float)x) : (T)trunc((double)x));
        case NDB_FN_SIGNUM: return x != x ? x : (T)copysign(1.0, (double)x);     // +-0 -> +-1, NaN -> NaN
        case NDB_FN_RECIP: return T(1) / x;
        case NDB_FN_EXP: return F ? (T)expf((float)x) : (T)exp((double)x);
        case NDB_FN_EXP2: return F ? (T)exp2f((float)x) : (T)exp2((double)x);
        case NDB_FN_EXP_M1: return F ? (T)expm1f((float)x) : (T)expm1((double)x);
        case NDB_FN_LN: return F ? (T)logf((float)x) : (T)log((double)x);
        case NDB_FN_LOG2: return F ? (T)log2f((float)x) : (T)log2((double)x);
        case NDB_FN_LOG10: return F ? (T)log10f((float)x) : (T)log10((double)x);
        case NDB_FN_LN_1P: return F ? (T)log1pf((float)x) : (T)log1p((double)x);
        case NDB_FN_CBRT: return F ? (T)cbrtf((float)x) : (T)cbrt((double)x);
        case NDB_FN_SIN: return F ? (T)sinf((float)x) : (T)sin((double)x);
        case NDB_FN_COS: return F ? (T)cosf((float)x) : (T)cos((double)x);
        case NDB_FN_TAN: return F ? (T)tanf((float)x) : (T)tan((double)x);
        case NDB_FN_ASIN: return F ? (T)asinf((float)x) : (T)asin((double)x);
        case NDB_FN_ACOS: return F ? (T)acosf((float)x) : (T)acos((double)x);
        case NDB_FN_ATAN: return F ? (T)atanf((float)x) : (T)atan((double)x);
        case NDB_FN_SINH: return F ? (T)sinhf((float)x) : (T)sinh((double)x);
        case NDB_FN_COSH: return F ? (T)coshf((float)x) : (T)cosh((double)x);
        case NDB_FN_TANH: return F ? (T)tanhf((float)x) : (T)tanh((double)x);
        case NDB_FN_ASINH: return F ? (T)asinhf((float)x) : (T)asinh((double)x);
        case NDB_FN_ACOSH: return F ? (T)acoshf((float)x) : (T)acosh((double)x);
        case NDB_FN_ATANH: return F ? (T)atanhf((float)x) : (T)atanh((double)x);
        case NDB_FN_TO_DEGREES: return x * (F ? (T)57.2957795130823208767981548141051703f : (T)(180.0 / 3.14159265358979323846264338327950288));
        case NDB_FN_TO_RADIANS: return x * (F ? (T)(3.14159265358979323846264338327950288f / 180.0f) : (T)(3.14159265358979323846264338327950288 / 180.0));
        case NDB_FN_POW2: return x * x;
        case NDB_FN_POWI: {  // compiler-rt's __powisf2 / __powidf2, which is what Rust's powi lowers to
            int b = (int)p;
            const bool recip = b < 0;
            T a = x, r = T(1);
            for (;;) {
                if (b & 1) r *= a;
                b /= 2;
                if (b == 0) break;
                a *= a;
            }
            return recip ? T(1) / r : r;
        }
        case NDB_FN_POWF: return F ? (T)powf((float)x, (float)p) : (T)pow((double)x, (double)p);
        case NDB_FN_LOG: return F ? (T)(logf((float)x) / logf((float)p)) : (T)(log((double)x) / log((double)p));
        case NDB_FN_ABS_SUB: return F ? (T)fdimf((float)x, (float)p) : (T)fdim((double)x, (double)p);
        case NDB_FN_HYPOT: return F ? (T)hypotf((float)x, (float)p) : (T)hypot((double)x, (double)p);
        case NDB_FN_IS_NAN: return x != x ? T(1) : T(0);
        case NDB_FN_IS_INFINITE: return (x == x && (x - x) != (x - x)) ? T(1) : T(0);
        default: return x;
        }
    } else {
        return x;
    }
}

// The op vocabulary (ndb_op).  `x` holds the NIN loaded inputs.
template <typename T, int OP, int NIN>
__device__ __forceinline__ T apply_op(const T *x, T s0, T s1, int *flag)
{
    if constexpr (OP == NDB_OP_COPY) return x[0];
    else if constexpr (OP == NDB_OP_FILL) return s0;
    else if constexpr (OP == NDB_OP_ADD) return w_add(x[0], x[1]);
    else if constexpr (OP == NDB_OP_SUB) return w_sub(x[0], x[1]);
    else if constexpr (OP == NDB_OP_MUL) return w_mul(x[0], x[1]);
    else if constexpr (OP == NDB_OP_DIV) return w_div(x[0], x[1], flag);
    else if constexpr (OP == NDB_OP_REM) return w_rem(x[0], x[1], flag);
    else if constexpr (OP == NDB_OP_MUL_ADD) return w_add(w_mul(x[0], x[1]), x[2]);
    else if constexpr (OP == NDB_OP_ADD_S) return w_add(x[0], s0);
    else if constexpr (OP == NDB_OP_SUB_S) return w_sub(x[0], s0);
    else if constexpr (OP == NDB_OP_MUL_S) return w_mul(x[0], s0);
    else if constexpr (OP == NDB_OP_DIV_S) return w_div(x[0], s0, flag);
    else if constexpr (OP == NDB_OP_REM_S) return w_rem(x[0], s0, flag);
    else if constexpr (OP == NDB_OP_RSUB_S) return w_sub(s0, x[0]);
    else if constexpr (OP == NDB_OP_RDIV_S) return w_div(s0, x[0], flag);
    else if constexpr (OP == NDB_OP_NEG) return w_neg(x[0]);
    else if constexpr (OP == NDB_OP_AXPY) return w_add(x[0], w_mul(s0, x[1]));
    else if constexpr (OP == NDB_OP_ADD_N) {
        T acc = w_add(x[0], x[1]);
#pragma unroll
        for (int i = 2; i < NIN; ++i) acc = w_add(acc, x[i]);
        return acc;
    } else if constexpr (OP == NDB_OP_ABS) return w_abs(x[0]);
    else if constexpr (OP == NDB_OP_SQRT) {
        if constexpr (std::is_same<T, float>::value) return sqrtf(x[0]);
        else if constexpr (std::is_same<T, double>::value) return sqrt(x[0]);
        else return x[0];
    } else if constexpr (OP == NDB_OP_MIN) return w_min(x[0], x[1]);
    else if constexpr (OP == NDB_OP_MAX) return w_max(x[0], x[1]);
    else if constexpr (OP == NDB_OP_SCALE_ADD) return w_add(w_mul(x[0], s0), w_mul(x[1], s1));
    else if constexpr (OP == NDB_OP_EQ) return x[0] == x[1] ? T(1) : T(0);
    else if constexpr (OP == NDB_OP_NE) return x[0] != x[1] ? T(1) : T(0);
    else if constexpr (OP == NDB_OP_LT) return x[0] < x[1] ? T(1) : T(0);
    else if constexpr (OP == NDB_OP_LE) return x[0] <= x[1] ? T(1) : T(0);
    else if constexpr (OP == NDB_OP_GT) return x[0] > x[1] ? T(1) : T(0);
    else if constexpr (OP == NDB_OP_GE) return x[0] >= x[1] ? T(1) : T(0);
    else if constexpr (OP == NDB_OP_MATH) return math_fn<T>((int)s0, x[0], s1);
    else if constexpr (OP == NDB_OP_CLAMP) return x[0] != x[0] ? x[0] : (x[0] < s0 ? s0 : (x[0] > s1 ? s1 : x[0]));
    else if constexpr (OP == NDB_OP_TF32_HI || OP == NDB_OP_TF32_LO) {
        if constexpr (std::is_same<T, float>::value) {
            uint32_t hi;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x[0]));
            const float h = __uint_as_float(hi);
            // non-finite values keep their hi part and get a zero lo part (inf - inf would poison the product)
            if constexpr (OP == NDB_OP_TF32_HI) return h;
            else return (fabsf(x[0]) <= 3.4028234664e38f) ? x[0] - h : 0.f;
        } else return x[0];
    } else return x[0];
}

// ---------------------------------------------------------------------------------------------
// flat, 128-bit vectorised
template <int NP>
struct FlatArgs {
    char *ptr[NP];
    int64_t n;      // elements
    int64_t n_vec;  // full 16-byte vectors
};

template <typename T, int OP, int NIN, int UNROLL>
__global__ void __launch_bounds__(256) zip_flat_vec(FlatArgs<NIN + 1> a, T s0, T s1, int *flag)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    constexpr int NI = NIN > 0 ? NIN : 1;
    const int64_t base = (int64_t)blockIdx.x * (256 * UNROLL) + threadIdx.x;
    V in[UNROLL][NI];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int64_t idx = base + (int64_t)u * 256;
        if (idx < a.n_vec) {
#pragma unroll
            for (int p = 0; p < NIN; ++p) in[u][p] = reinterpret_cast<const V *>(a.ptr[p + 1])[idx];
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int64_t idx = base + (int64_t)u * 256;
        if (idx < a.n_vec) {
            V o;
#pragma unroll
            for (int e = 0; e < N; ++e) {
                T x[NI];
#pragma unroll
                for (int p = 0; p < NIN; ++p) x[p] = in[u][p].v[e];
                o.v[e] = apply_op<T, OP, NIN>(x, s0, s1, flag);
            }
            reinterpret_cast<V *>(a.ptr[0])[idx] = o;
        }
    }
    // ragged tail (< N elements), one thread each
    if (blockIdx.x == 0) {
        const int64_t i = a.n_vec * N + threadIdx.x;
        if (i < a.n) {
            T x[NI];
#pragma unroll
            for (int p = 0; p < NIN; ++p) x[p] = reinterpret_cast<const T *>(a.ptr[p + 1])[i];
            reinterpret_cast<T *>(a.ptr[0])[i] = apply_op<T, OP, NIN>(x, s0, s1, flag);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// collapsed N-d rows
//
// Geometry: [batch axes...] x [row axis] x [inner axis].  The inner axis is the output's fastest axis; the
// row axis is the longest remaining axis; whatever is left is the batch, whose multi-index is decomposed
// once per BLOCK (uniform work), never per thread.  A 256-thread block covers (TY * RI) rows x (TX * CI)
// units and every thread owns RI x CI = ITEMS units (CI chosen on the host so that short rows do not leave
// items idle).  Phase 1 computes every address and issues every load of the thread with no branch in
// between; phase 2 applies the op and stores.
//   VEC   128-bit units.
//   PLAIN every producer runs forwards along the row (stride 1).  Otherwise an input may also be constant
//         along the row (stride 0: one scalar load) or run BACKWARDS (stride -1: the mirrored vector,
//         lanes reversed in registers) - predicated loads, still no branch, so s![.., ..;-1] views and
//         row-broadcast operands stay on the vector path.
template <int NP>
struct RowsArgs {
    char *ptr[NP];
    int64_t bstride[NP][NDB_MAX_NDIM - 2];  // batch-axis strides, elements
    int64_t bdim[NDB_MAX_NDIM - 2];
    int n_batch;
    int64_t rstride[NP];  // row-axis stride, elements
    int64_t istride[NP];  // inner stride, elements (VEC: 1, 0 or -1)
    int64_t n_rows;
    int64_t inner_len;  // in units (elements, or vectors when VEC)
    int64_t blocks_inner, blocks_rows, total_blocks;
};

template <typename T, int OP, int NIN, bool VEC, bool PLAIN, int ITEMS>
__global__ void __launch_bounds__(256, (NIN <= 3 ? 3 : 1)) zip_rows(RowsArgs<NIN + 1> a, int ci_shift, T s0, T s1, int *flag)
{
    pdl_enter();
    using V = Vec128<T>;
    using E = typename std::conditional<VEC, V, T>::type;  // the unit moved per access
    constexpr int N = VEC ? V::N : 1;
    constexpr int NI = NIN > 0 ? NIN : 1;
    constexpr int NP = NIN + 1;
    const int TX = blockDim.x, TY = blockDim.y;
    const int CI = 1 << ci_shift, RI = ITEMS >> ci_shift;
    for (int64_t vb = blockIdx.x; vb < a.total_blocks; vb += gridDim.x) {
        // ---- per block (uniform): which batch, which rows, which part of the row
        int64_t bin = 0, brow = 0;
        const T *base[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) base[p] = reinterpret_cast<const T *>(a.ptr[p]);
        if (a.total_blocks <= 0x7fffffffLL) {  // the usual case: 32-bit divisions
            unsigned t = (unsigned)vb;
            if (a.blocks_inner > 1) {
                const unsigned q = t / (unsigned)a.blocks_inner;
                bin = t - q * (unsigned)a.blocks_inner;
                t = q;
            }
            if (a.n_batch > 0) {
                const unsigned q = t / (unsigned)a.blocks_rows;
                brow = t - q * (unsigned)a.blocks_rows;
                t = q;
            } else {
                brow = t;
            }
#pragma unroll 1
            for (int k = a.n_batch - 1; k >= 0; --k) {
                const unsigned q = t / (unsigned)a.bdim[k];
                const int64_t idx = (int64_t)(t - q * (unsigned)a.bdim[k]);
                t = q;
#pragma unroll
                for (int p = 0; p < NP; ++p) base[p] += idx * a.bstride[p][k];
            }
        } else {
            int64_t t = vb;
            if (a.blocks_inner > 1) {
                const int64_t q = t / a.blocks_inner;
                bin = t - q * a.blocks_inner;
                t = q;
            }
            if (a.n_batch > 0) {
                const int64_t q = t / a.blocks_rows;
                brow = t - q * a.blocks_rows;
                t = q;
            } else {
                brow = t;
            }
#pragma unroll 1
            for (int k = a.n_batch - 1; k >= 0; --k) {
                const int64_t q = t / a.bdim[k];
                const int64_t idx = t - q * a.bdim[k];
                t = q;
#pragma unroll
                for (int p = 0; p < NP; ++p) base[p] += idx * a.bstride[p][k];
            }
        }
        const int64_t row0 = brow * ((int64_t)TY * RI) + threadIdx.y;
        const int64_t i0 = bin * ((int64_t)TX * CI) + threadIdx.x;
        E in[ITEMS][NI];
        T *dst[ITEMS];
        bool live[ITEMS];
        // ---- phase 1: addresses + loads
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int r_it = it >> ci_shift, c_it = it & (CI - 1);
            const int64_t row = row0 + (int64_t)r_it * TY;
            const int64_t i = i0 + (int64_t)c_it * TX;
            live[it] = row < a.n_rows && i < a.inner_len;
            if constexpr (PLAIN) {
                dst[it] = const_cast<T *>(base[0]) + row * a.rstride[0] + i * N;
#pragma unroll
                for (int p = 0; p < NIN; ++p) {
                    const T *src = base[p + 1] + row * a.rstride[p + 1] + i * N;
                    if (live[it]) in[it][p] = *reinterpret_cast<const E *>(src);
                }
            } else {
                dst[it] = const_cast<T *>(base[0]) + row * a.rstride[0] + i * N * a.istride[0];
#pragma unroll
                for (int p = 0; p < NIN; ++p) {
                    const int64_t is = a.istride[p + 1];
                    if constexpr (VEC) {
                        const T *src = base[p + 1] + row * a.rstride[p + 1] + i * N * is - (is < 0 ? (N - 1) : 0);
                        if (live[it] && is != 0) in[it][p] = *reinterpret_cast<const V *>(src);
                        if (live[it] && is == 0) in[it][p].v[0] = *src;
                    } else {
                        if (live[it]) in[it][p] = (base[p + 1] + row * a.rstride[p + 1])[i * is];
                    }
                }
            }
        }
        // ---- phase 2: op + store
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            if (!live[it]) continue;
            if constexpr (VEC) {
                V o;
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    T x[NI];
#pragma unroll
                    for (int p = 0; p < NIN; ++p) {
                        if constexpr (PLAIN) x[p] = in[it][p].v[e];
                        else {
                            const int64_t is = a.istride[p + 1];
                            x[p] = is > 0 ? in[it][p].v[e] : (is < 0 ? in[it][p].v[N - 1 - e] : in[it][p].v[0]);
                        }
                    }
                    o.v[e] = apply_op<T, OP, NIN>(x, s0, s1, flag);
                }
                *reinterpret_cast<V *>(dst[it]) = o;
            } else {
                *dst[it] = apply_op<T, OP, NIN>(in[it], s0, s1, flag);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// shared-memory tile transpose
template <int NP>
struct TileArgs {
    char *ptr[NP];
    int64_t xs[NP], ys[NP];  // strides along X (output's fastest axis) and Y (staged axis), elements
    int64_t ostride[NP][NDB_MAX_NDIM - 2];
    int64_t odim[NDB_MAX_NDIM - 2];
    int n_outer;
    int64_t nx, ny, tiles_x, tiles_y, total_tiles;
    int slot[NP];   // shared-memory slot of a staged producer, -1 = read straight from global
    int y_fastest;  // block order: 1 = consecutive blocks walk Y tiles first
};

#ifndef NDB_TILE_MINB_STEP2
#define NDB_TILE_MINB_STEP2 2
#endif
#ifndef NDB_TILE_MINB3
#define NDB_TILE_MINB3 2
#endif
#ifndef NDB_TILE_MINB2
#define NDB_TILE_MINB2 4  // CTAs/SM the binary-op tile kernel is compiled for (64 registers)
#endif

template <typename T>
struct TileShape {
    static constexpr int TS = sizeof(T) == 4 ? 64 : 32;  // 256-byte runs either way
};

template <int NP>
__device__ __forceinline__ void tile_coords(const TileArgs<NP> &a, int64_t vb, int64_t &tx, int64_t &ty, int64_t (&off)[NP])
{
#pragma unroll
    for (int p = 0; p < NP; ++p) off[p] = 0;
    if (a.total_tiles <= 0x7fffffff) {
        // the usual case: 32-bit divisions (a 64-bit division is ~10x the instructions, and this runs once per tile)
        uint32_t t = (uint32_t)vb;
        const uint32_t f = (uint32_t)(a.y_fastest ? a.tiles_y : a.tiles_x), sl = (uint32_t)(a.y_fastest ? a.tiles_x : a.tiles_y);
        const uint32_t fast = t % f;
        t /= f;
        const uint32_t slow = t % sl;
        t /= sl;
        tx = a.y_fastest ? slow : fast;
        ty = a.y_fastest ? fast : slow;
        for (int k = a.n_outer - 1; k >= 0; --k) {
            const uint32_t d = (uint32_t)a.odim[k];
            const uint32_t q = t / d;
            const uint32_t idx = t - q * d;
            t = q;
#pragma unroll
            for (int p = 0; p < NP; ++p) off[p] += (int64_t)idx * a.ostride[p][k];
        }
        return;
    }
    int64_t t = vb;
    if (a.y_fastest) {
        ty = t % a.tiles_y;
        t /= a.tiles_y;
        tx = t % a.tiles_x;
        t /= a.tiles_x;
    } else {
        tx = t % a.tiles_x;
        t /= a.tiles_x;
        ty = t % a.tiles_y;
        t /= a.tiles_y;
    }
    for (int k = a.n_outer - 1; k >= 0; --k) {
        const int64_t q = t / a.odim[k];
        const int64_t idx = t - q * a.odim[k];
        t = q;
#pragma unroll
        for (int p = 0; p < NP; ++p) off[p] += idx * a.ostride[p][k];
    }
}

// scalar form: any signed strides
template <typename T, int OP, int NIN>
__global__ void __launch_bounds__(256) zip_tile(TileArgs<NIN + 1> a, T s0, T s1, int *flag)
{
    pdl_enter();
    constexpr int TS = TileShape<T>::TS;
    constexpr int PASS = 256 / TS;  // rows (or columns) covered per pass
    constexpr int NI = NIN > 0 ? NIN : 1;
    constexpr int NP = NIN + 1;
    extern __shared__ __align__(16) unsigned char tile_raw[];  // one padded TS x (TS+1) tile per staged producer
    T(*tile)[TS][TS + 1] = reinterpret_cast<T(*)[TS][TS + 1]>(tile_raw);
    const int lo = threadIdx.x % TS;  // fast lane index
    const int hi = threadIdx.x / TS;
    for (int64_t vb = blockIdx.x; vb < a.total_tiles; vb += gridDim.x) {
        int64_t tx, ty, off[NP];
        tile_coords<NP>(a, vb, tx, ty, off);
        const int64_t x0 = tx * TS, y0 = ty * TS;
        // stage: lanes run along Y (the staged producer's dense axis)
#pragma unroll
        for (int p = 1; p < NP; ++p) {
            const int sl = a.slot[p];
            if (sl >= 0) {
                const T *src = reinterpret_cast<const T *>(a.ptr[p]) + off[p];
                const int64_t y = y0 + lo;
#pragma unroll 4
                for (int ps = 0; ps < TS / PASS; ++ps) {
                    const int lx = hi + ps * PASS;
                    const int64_t x = x0 + lx;
                    if (y < a.ny && x < a.nx) tile[sl][lx][lo] = src[y * a.ys[p] + x * a.xs[p]];
                }
            }
        }
        __syncthreads();
        // compute: lanes run along X
        {
            const int64_t x = x0 + lo;
            T *dst = reinterpret_cast<T *>(a.ptr[0]) + off[0];
#pragma unroll 4
            for (int ps = 0; ps < TS / PASS; ++ps) {
                const int ly = hi + ps * PASS;
                const int64_t y = y0 + ly;
                if (y < a.ny && x < a.nx) {
                    T v[NI];
#pragma unroll
                    for (int p = 1; p < NP; ++p) {
                        const int sl = a.slot[p];
                        if (sl >= 0) v[p - 1] = tile[sl][lo][ly];
                        else v[p - 1] = (reinterpret_cast<const T *>(a.ptr[p]) + off[p])[y * a.ys[p] + x * a.xs[p]];
                    }
                    dst[y * a.ys[0] + x * a.xs[0]] = apply_op<T, OP, NIN>(v, s0, s1, flag);
                }
            }
        }
        __syncthreads();
    }
}

// 128-bit form: staged producers have ys == 1, everything else xs == 1, all rows 16-byte aligned and the
// tile extents multiples of the vector width (checked on the host).  Every thread first issues ALL its
// global loads (PASSES vectors per producer), then the staged ones go through shared memory.
// STEP2: a staged producer may also step by +-2 along its dense axis (x.t().slice(s![..;-2, ..])): two vectors are
// loaded and every other lane kept.
// FULL = the tile lies entirely inside the array (every tile but the last row/column of tiles): bounds predicates fold away
template <typename T, int OP, int NIN, bool STEP2, bool FULL>
__device__ __forceinline__ void zip_tile_vec_body(const TileArgs<NIN + 1> &a, T s0, T s1, int *flag, unsigned char *tile_raw, int64_t tx,
                                                  int64_t ty, const int64_t (&off)[NIN + 1])
{
    using V = Vec128<T>;
    constexpr int N = V::N;
    constexpr int TS = TileShape<T>::TS;
    constexpr int VPR = TS / N;       // vectors per tile row (16)
    constexpr int RPP = 256 / VPR;    // tile rows per pass (16)
    constexpr int PASSES = TS / RPP;  // 4 (f32) / 2 (f64)
    constexpr int NI = NIN > 0 ? NIN : 1;
    constexpr int NP = NIN + 1;
    T(*tile)[TS][TS + 1] = reinterpret_cast<T(*)[TS][TS + 1]>(tile_raw);
    const int vl = threadIdx.x % VPR;
    const int hi = threadIdx.x / VPR;
    {
        const int64_t x0 = tx * TS, y0 = ty * TS;
        V ld[NI][PASSES];
        V ld2[STEP2 ? NI : 1][STEP2 ? PASSES : 1];  // second vector of a +-2-stepped staged producer
#pragma unroll
        for (int p = 1; p < NP; ++p) {
            const T *src = reinterpret_cast<const T *>(a.ptr[p]) + off[p];
            if (a.slot[p] >= 0) {
                // lanes along Y.  The N logical elements y..y+N-1 live in |ys| * N consecutive memory elements starting
                // at A (a backwards axis is read as the mirrored vectors); vectors are aligned by the host's rules.
                const int64_t y = y0 + (int64_t)vl * N;
                const int64_t sy = a.ys[p] < 0 ? -a.ys[p] : a.ys[p];
                const int64_t A = a.ys[p] > 0 ? y * sy : -(y * sy) - (sy * N - 1);
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int64_t x = x0 + hi + ps * RPP;
                    if (FULL || (x < a.nx && y < a.ny)) {
                        ld[p - 1][ps] = *reinterpret_cast<const V *>(src + A + x * a.xs[p]);
                        if constexpr (STEP2) {
                            if (sy == 2) ld2[p - 1][ps] = *reinterpret_cast<const V *>(src + A + N + x * a.xs[p]);
                        }
                    }
                }
            } else {
                const int64_t x = x0 + (int64_t)vl * N;  // lanes along X
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int64_t y = y0 + hi + ps * RPP;
                    if (FULL || (x < a.nx && y < a.ny)) ld[p - 1][ps] = *reinterpret_cast<const V *>(src + y * a.ys[p] + x);
                }
            }
        }
#pragma unroll
        for (int p = 1; p < NP; ++p) {
            const int sl = a.slot[p];
            if (sl >= 0) {
                const int64_t ys = a.ys[p];
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int lx = hi + ps * RPP;
                    if (FULL || (x0 + lx < a.nx && y0 + (int64_t)vl * N < a.ny)) {
                        const V &v0 = ld[p - 1][ps];
#pragma unroll
                        for (int e = 0; e < N; ++e) {
                            T val = ys > 0 ? v0.v[e] : v0.v[N - 1 - e];  // |ys| == 1
                            if constexpr (STEP2) {
                                const V &v1 = ld2[p - 1][ps];
                                constexpr int NN = N;
                                const T fwd = (2 * e < NN) ? v0.v[(2 * e) % NN] : v1.v[(2 * e) % NN];
                                const T bwd = (2 * NN - 1 - 2 * e < NN) ? v0.v[(2 * NN - 1 - 2 * e) % NN] : v1.v[(2 * NN - 1 - 2 * e) % NN];
                                if (ys == 2) val = fwd;
                                if (ys == -2) val = bwd;
                            }
                            tile[sl][lx][vl * N + e] = val;
                        }
                    }
                }
            }
        }
        __syncthreads();
        {
            const int64_t x = x0 + (int64_t)vl * N;
            T *dst = reinterpret_cast<T *>(a.ptr[0]) + off[0];
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) {
                const int ly = hi + ps * RPP;
                const int64_t y = y0 + ly;
                if (FULL || (x < a.nx && y < a.ny)) {
                    V o;
#pragma unroll
                    for (int e = 0; e < N; ++e) {
                        T v[NI];
#pragma unroll
                        for (int p = 1; p < NP; ++p) {
                            const int sl = a.slot[p];
                            v[p - 1] = sl >= 0 ? tile[sl][vl * N + e][ly] : ld[p - 1][ps].v[e];
                        }
                        o.v[e] = apply_op<T, OP, NIN>(v, s0, s1, flag);
                    }
                    *reinterpret_cast<V *>(dst + y * a.ys[0] + x) = o;
                }
            }
        }
        __syncthreads();
    }
}

// +-2-stepped staged producer, the common case of exactly ONE staged producer whose index SP is known at compile time: the
// generic STEP2 body keeps first AND second vectors for every producer because which one is staged is a run-time fact
// (96 data registers, 128 in all, 2 CTAs/SM); with SP fixed only the staged producer carries a second vector and only the
// others carry direct ones (64 data registers), which fits the 3-CTA budget.
template <typename T, int OP, int NIN, int SP, bool FULL>
__device__ __forceinline__ void zip_tile_s2_body(const TileArgs<NIN + 1> &a, T s0, T s1, int *flag, unsigned char *tile_raw, int64_t tx,
                                                 int64_t ty, const int64_t (&off)[NIN + 1])
{
    using V = Vec128<T>;
    constexpr int N = V::N;
    constexpr int TS = TileShape<T>::TS;
    constexpr int VPR = TS / N;
    constexpr int RPP = 256 / VPR;
    constexpr int PASSES = TS / RPP;
    constexpr int NI = NIN > 0 ? NIN : 1;
    constexpr int NP = NIN + 1;
    T(*tile)[TS + 1] = reinterpret_cast<T(*)[TS + 1]>(tile_raw);  // one staged producer: one tile
    const int vl = threadIdx.x % VPR;
    const int hi = threadIdx.x / VPR;
    const int64_t x0 = tx * TS, y0 = ty * TS;
    V st0[PASSES], st1[PASSES];
    const int64_t ys = a.ys[SP];
    {
        const T *src = reinterpret_cast<const T *>(a.ptr[SP]) + off[SP];
        const int64_t y = y0 + (int64_t)vl * N;
        const int64_t sy = ys < 0 ? -ys : ys;
        const int64_t A = ys > 0 ? y * sy : -(y * sy) - (sy * N - 1);
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
            const int64_t x = x0 + hi + ps * RPP;
            if (FULL || (x < a.nx && y < a.ny)) {
                st0[ps] = *reinterpret_cast<const V *>(src + A + x * a.xs[SP]);
                if (sy == 2) st1[ps] = *reinterpret_cast<const V *>(src + A + N + x * a.xs[SP]);
            }
        }
    }
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
        const int lx = hi + ps * RPP;
        if (FULL || (x0 + lx < a.nx && y0 + (int64_t)vl * N < a.ny)) {
#pragma unroll
            for (int e = 0; e < N; ++e) {
                T val = ys > 0 ? st0[ps].v[e] : st0[ps].v[N - 1 - e];  // |ys| == 1
                const T fwd = (2 * e < N) ? st0[ps].v[(2 * e) % N] : st1[ps].v[(2 * e) % N];
                const T bwd = (2 * N - 1 - 2 * e < N) ? st0[ps].v[(2 * N - 1 - 2 * e) % N] : st1[ps].v[(2 * N - 1 - 2 * e) % N];
                if (ys == 2) val = fwd;
                if (ys == -2) val = bwd;
                tile[lx][vl * N + e] = val;
            }
        }
    }
    // Phase 2 walks the tile rows in groups of GP passes: the direct producers' vectors of a group are loaded right before they
    // are consumed (the first group's loads are in flight across the barrier).  Three inputs: two passes per group, which
    // keeps the kernel inside the 3-CTA register budget without spills; fewer inputs: all passes at once.
    constexpr int GP = (NIN >= 3 && PASSES >= 2) ? PASSES / 2 : PASSES;  // (three-input instantiations are not dispatched today)
    const int64_t xg = x0 + (int64_t)vl * N;
    T *dst = reinterpret_cast<T *>(a.ptr[0]) + off[0];
#pragma unroll
    for (int g0 = 0; g0 < PASSES; g0 += GP) {
        V ld[NI][GP];  // direct producers (entry SP - 1 stays unused and is dropped by the compiler)
#pragma unroll
        for (int p = 1; p < NP; ++p) {
            if (p == SP) continue;
            const T *src = reinterpret_cast<const T *>(a.ptr[p]) + off[p];
#pragma unroll
            for (int q = 0; q < GP; ++q) {
                const int64_t y = y0 + hi + (g0 + q) * RPP;
                if (FULL || (xg < a.nx && y < a.ny)) ld[p - 1][q] = *reinterpret_cast<const V *>(src + y * a.ys[p] + xg);
            }
        }
        if (g0 == 0) __syncthreads();  // the staged tile is complete
#pragma unroll
        for (int q = 0; q < GP; ++q) {
            const int ly = hi + (g0 + q) * RPP;
            const int64_t y = y0 + ly;
            if (FULL || (xg < a.nx && y < a.ny)) {
                V o;
#pragma unroll
                for (int e = 0; e < N; ++e) {
                    T v[NI];
#pragma unroll
                    for (int p = 1; p < NP; ++p) v[p - 1] = p == SP ? tile[vl * N + e][ly] : ld[p - 1][q].v[e];
                    o.v[e] = apply_op<T, OP, NIN>(v, s0, s1, flag);
                }
                *reinterpret_cast<V *>(dst + y * a.ys[0] + xg) = o;
            }
        }
    }
    __syncthreads();
}

template <typename T, int OP, int NIN, int SP>
__global__ void __launch_bounds__(256, 3) zip_tile_vec_s2(TileArgs<NIN + 1> a, T s0, T s1, int *flag)
{
    pdl_enter();
    constexpr int TS = TileShape<T>::TS;
    extern __shared__ __align__(16) unsigned char tile_raw[];
    const int64_t full_x = a.nx / TS, full_y = a.ny / TS;
    for (int64_t vb = blockIdx.x; vb < a.total_tiles; vb += gridDim.x) {
        int64_t tx, ty, off[NIN + 1];
        tile_coords<NIN + 1>(a, vb, tx, ty, off);
        if (tx < full_x && ty < full_y) zip_tile_s2_body<T, OP, NIN, SP, true>(a, s0, s1, flag, tile_raw, tx, ty, off);
        else zip_tile_s2_body<T, OP, NIN, SP, false>(a, s0, s1, flag, tile_raw, tx, ty, off);
    }
}

template <typename T, int OP, int NIN, bool STEP2>
__global__ void __launch_bounds__(256, STEP2 ? NDB_TILE_MINB_STEP2 : (NIN <= 1 ? 6 : (NIN <= 2 ? NDB_TILE_MINB2 : NDB_TILE_MINB3))) zip_tile_vec(TileArgs<NIN + 1> a, T s0, T s1, int *flag)
{
    pdl_enter();
    constexpr int TS = TileShape<T>::TS;
    extern __shared__ __align__(16) unsigned char tile_raw[];  // one padded TS x (TS+1) tile per staged producer
    const int64_t full_x = a.nx / TS, full_y = a.ny / TS;
    for (int64_t vb = blockIdx.x; vb < a.total_tiles; vb += gridDim.x) {
        int64_t tx, ty, off[NIN + 1];
        tile_coords<NIN + 1>(a, vb, tx, ty, off);
        if (tx < full_x && ty < full_y) zip_tile_vec_body<T, OP, NIN, STEP2, true>(a, s0, s1, flag, tile_raw, tx, ty, off);
        else zip_tile_vec_body<T, OP, NIN, STEP2, false>(a, s0, s1, flag, tile_raw, tx, ty, off);
    }
}

inline int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---------------------------------------------------------------------------------------------
// host-side launch
template <typename T, int OP, int NIN>
int launch_zip(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    constexpr int NP = NIN + 1;
    T s0 = T(0), s1 = T(0);
    const int ns = op_scalars(OP);
    if (ns > 0) {
        if (!scalars) {
            set_error("Zip: op %d needs %d scalar(s)", OP, ns);
            return NDB_ERR_INVALID_ARG;
        }
        s0 = reinterpret_cast<const T *>(scalars)[0];
        if (ns > 1) s1 = reinterpret_cast<const T *>(scalars)[1];
    }
    cudaStream_t st = ctx->stream;
    const int m = z.ndim;
    constexpr int64_t VN = Vec128<T>::N;
    if (z.kernel == NDB_KERNEL_FLAT_VEC) {
        constexpr int UNROLL = NIN <= 3 ? 4 : (NIN <= 5 ? 2 : 1);
        FlatArgs<NP> a;
        for (int p = 0; p < NP; ++p) a.ptr[p] = z.ptr[p];
        a.n = z.total;
        a.n_vec = z.total / VN;
        int64_t blocks = (a.n_vec + 256 * UNROLL - 1) / (256 * UNROLL);
        if (blocks < 1) blocks = 1;
        if (blocks > 0x7fffffffLL) {
            set_error("Zip: %lld elements exceed one launch", (long long)z.total);
            return NDB_ERR_UNSUPPORTED;
        }
        launch_pdl(zip_flat_vec<T, OP, NIN, UNROLL>, (unsigned)blocks, 256, 0, st, a, s0, s1, ctx->dev_flag);
        NDB_CHECK_LAUNCH();
        return NDB_OK;
    }
    if (z.kernel == NDB_KERNEL_TILE) {
        constexpr int TS = TileShape<T>::TS;
        TileArgs<NP> a;
        memset(&a, 0, sizeof(a));
        const int X = m - 1, Y = z.tile_axis;
        int nslot = 0;
        for (int p = 0; p < NP; ++p) {
            a.ptr[p] = z.ptr[p];
            a.xs[p] = z.stride[p][X];
            a.ys[p] = z.stride[p][Y];
            a.slot[p] = (z.staged_mask >> p) & 1 ? nslot++ : -1;
        }
        int no = 0;
        for (int k = 0; k < m; ++k) {
            if (k == X || k == Y) continue;
            a.odim[no] = z.dim[k];
            for (int p = 0; p < NP; ++p) a.ostride[p][no] = z.stride[p][k];
            ++no;
        }
        a.n_outer = no;
        a.nx = z.dim[X];
        a.ny = z.dim[Y];
        a.tiles_x = (a.nx + TS - 1) / TS;
        a.tiles_y = (a.ny + TS - 1) / TS;
        a.total_tiles = a.tiles_x * a.tiles_y;
        for (int k = 0; k < no; ++k) a.total_tiles *= a.odim[k];
        a.y_fastest = env_int("NDB_TILE_YFAST", 0);
        // vector form?
        bool vec = a.nx % VN == 0 && a.ny % VN == 0, step2 = false;
        for (int p = 0; p < NP && vec; ++p) {
            if (a.slot[p] >= 0) {
                // forwards: base aligned; backwards: the element AFTER the base is aligned (mirrored vectors)
                const int64_t sy = llabs(a.ys[p]);
                vec = a.xs[p] % VN == 0 && (sy == 1 || sy == 2) &&
                      ((a.ys[p] > 0 && ((uintptr_t)a.ptr[p] % 16) == 0) || (a.ys[p] < 0 && ((uintptr_t)(a.ptr[p] + sizeof(T)) % 16) == 0));
                step2 = step2 || sy == 2;
            } else {
                vec = ((uintptr_t)a.ptr[p] % 16) == 0 && a.xs[p] == 1 && a.ys[p] % VN == 0;
            }
            for (int k = 0; k < no && vec; ++k) vec = a.ostride[p][k] % VN == 0;
        }
        if (env_int("NDB_TILE_SCALAR", 0)) vec = false;
        int64_t blocks = a.total_tiles < (int64_t)1 << 30 ? a.total_tiles : (int64_t)1 << 30;
        const size_t smem = (size_t)nslot * TS * (TS + 1) * sizeof(T);
        if (vec && step2 && nslot == 1 && NIN >= 1 && NIN <= 2 && !env_int("NDB_TILE_S2_GENERIC", 0)) {
            // exactly one staged producer and at most two inputs: the kernel specialised on WHICH one is staged (fewer live
            // vectors, 3 CTAs/SM): x.t()[::-2] + y[::2] 4005 -> 4264 GB/s, x.t()[::2] copy 3352 -> 3788 GB/s.  With three
            // inputs the specialised form measured SLOWER than the generic one (3290 vs 3702 GB/s), so those keep it.
            int sp = 0;
            for (int p = 1; p < NP; ++p)
                if (a.slot[p] >= 0) sp = p;
            if constexpr (NIN >= 1) {
                if (sp == 1) launch_pdl(zip_tile_vec_s2<T, OP, NIN, 1>, (unsigned)blocks, 256, smem, st, a, s0, s1, ctx->dev_flag);
            }
            if constexpr (NIN >= 2) {
                if (sp == 2) launch_pdl(zip_tile_vec_s2<T, OP, NIN, 2>, (unsigned)blocks, 256, smem, st, a, s0, s1, ctx->dev_flag);
            }
        } else if (vec && step2) launch_pdl(zip_tile_vec<T, OP, NIN, true>, (unsigned)blocks, 256, smem, st, a, s0, s1, ctx->dev_flag);
        else if (vec) launch_pdl(zip_tile_vec<T, OP, NIN, false>, (unsigned)blocks, 256, smem, st, a, s0, s1, ctx->dev_flag);
        else launch_pdl(zip_tile<T, OP, NIN>, (unsigned)blocks, 256, smem, st, a, s0, s1, ctx->dev_flag);
        NDB_CHECK_LAUNCH();
        return NDB_OK;
    }
    // rows (FLAT with unaligned pointers is the 1-d case of ROWS)
    {
        const bool vec = z.kernel == NDB_KERNEL_ROWS_VEC;
        constexpr int ITEMS = NIN <= 3 ? 4 : 2;
        RowsArgs<NP> a;
        memset(&a, 0, sizeof(a));
        const int inner = m - 1;
        const int64_t unit = vec ? VN : 1;
        // the longest outer axis becomes the row axis, the others the batch
        int row_axis = -1;
        for (int k = 0; k < inner; ++k)
            if (row_axis < 0 || z.dim[k] > z.dim[row_axis]) row_axis = k;
        bool plain = true;
        for (int p = 0; p < NP; ++p) {
            a.ptr[p] = z.ptr[p];
            a.istride[p] = z.stride[p][inner];
            a.rstride[p] = row_axis >= 0 ? z.stride[p][row_axis] : 0;
            plain = plain && a.istride[p] == 1;
        }
        a.n_rows = row_axis >= 0 ? z.dim[row_axis] : 1;
        int nb = 0;
        int64_t batches = 1;
        for (int k = 0; k < inner; ++k) {
            if (k == row_axis) continue;
            a.bdim[nb] = z.dim[k];
            for (int p = 0; p < NP; ++p) a.bstride[p][nb] = z.stride[p][k];
            batches *= z.dim[k];
            ++nb;
        }
        a.n_batch = nb;
        a.inner_len = z.dim[inner] / unit;
        int tx = 256;
        while (tx > 1 && tx / 2 >= a.inner_len) tx /= 2;  // smallest power of two covering a short row
        if (tx < 1) tx = 1;
        int ty = 256 / tx;
        while (ty > 1 && ty / 2 >= a.n_rows) ty /= 2;  // few rows: a smaller block beats idle threads
        // per-thread items: along the row while it is long enough, the rest across rows
        int ci_shift = 0;
        while ((1 << ci_shift) < ITEMS && (int64_t)tx * (1 << ci_shift) < a.inner_len) ++ci_shift;
        const int ci = 1 << ci_shift, ri = ITEMS >> ci_shift;
        a.blocks_inner = (a.inner_len + (int64_t)tx * ci - 1) / ((int64_t)tx * ci);
        a.blocks_rows = (a.n_rows + (int64_t)ty * ri - 1) / ((int64_t)ty * ri);
        a.total_blocks = a.blocks_inner * a.blocks_rows * batches;
        int64_t blocks = a.total_blocks < (int64_t)1 << 30 ? a.total_blocks : (int64_t)1 << 30;
        dim3 bd(tx, ty, 1);
        const unsigned gb = (unsigned)blocks;
        if (vec && plain) launch_pdl(zip_rows<T, OP, NIN, true, true, ITEMS>, gb, bd, 0, st, a, ci_shift, s0, s1, ctx->dev_flag);
        else if (vec) launch_pdl(zip_rows<T, OP, NIN, true, false, ITEMS>, gb, bd, 0, st, a, ci_shift, s0, s1, ctx->dev_flag);
        else launch_pdl(zip_rows<T, OP, NIN, false, false, ITEMS>, gb, bd, 0, st, a, ci_shift, s0, s1, ctx->dev_flag);
        NDB_CHECK_LAUNCH();
        return NDB_OK;
    }
}

template <typename T>
int zip_dispatch_typed(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    const int nin = z.n_parts - 1;
    switch (z.op) {
#define NDB_ZIP_CASE(OP, NIN) \
    case OP: return launch_zip<T, OP, NIN>(ctx, z, scalars);
        NDB_ZIP_CASE(NDB_OP_COPY, 1)
        NDB_ZIP_CASE(NDB_OP_FILL, 0)
        NDB_ZIP_CASE(NDB_OP_ADD, 2)
        NDB_ZIP_CASE(NDB_OP_SUB, 2)
        NDB_ZIP_CASE(NDB_OP_MUL, 2)
        NDB_ZIP_CASE(NDB_OP_DIV, 2)
        NDB_ZIP_CASE(NDB_OP_REM, 2)
        NDB_ZIP_CASE(NDB_OP_MUL_ADD, 3)
        NDB_ZIP_CASE(NDB_OP_ADD_S, 1)
        NDB_ZIP_CASE(NDB_OP_SUB_S, 1)
        NDB_ZIP_CASE(NDB_OP_MUL_S, 1)
        NDB_ZIP_CASE(NDB_OP_DIV_S, 1)
        NDB_ZIP_CASE(NDB_OP_REM_S, 1)
        NDB_ZIP_CASE(NDB_OP_RSUB_S, 1)
        NDB_ZIP_CASE(NDB_OP_RDIV_S, 1)
        NDB_ZIP_CASE(NDB_OP_NEG, 1)
        NDB_ZIP_CASE(NDB_OP_AXPY, 2)
        NDB_ZIP_CASE(NDB_OP_ABS, 1)
        NDB_ZIP_CASE(NDB_OP_SQRT, 1)
        NDB_ZIP_CASE(NDB_OP_MIN, 2)
        NDB_ZIP_CASE(NDB_OP_MAX, 2)
        NDB_ZIP_CASE(NDB_OP_SCALE_ADD, 2)
        NDB_ZIP_CASE(NDB_OP_TF32_HI, 1)
        NDB_ZIP_CASE(NDB_OP_TF32_LO, 1)
        NDB_ZIP_CASE(NDB_OP_EQ, 2)
        NDB_ZIP_CASE(NDB_OP_NE, 2)
        NDB_ZIP_CASE(NDB_OP_LT, 2)
        NDB_ZIP_CASE(NDB_OP_LE, 2)
        NDB_ZIP_CASE(NDB_OP_GT, 2)
        NDB_ZIP_CASE(NDB_OP_GE, 2)
        NDB_ZIP_CASE(NDB_OP_MATH, 1)
        NDB_ZIP_CASE(NDB_OP_CLAMP, 1)
#undef NDB_ZIP_CASE
    case NDB_OP_ADD_N:
        switch (nin) {
        case 2: return launch_zip<T, NDB_OP_ADD_N, 2>(ctx, z, scalars);
        case 3: return launch_zip<T, NDB_OP_ADD_N, 3>(ctx, z, scalars);
        case 4: return launch_zip<T, NDB_OP_ADD_N, 4>(ctx, z, scalars);
        case 5: return launch_zip<T, NDB_OP_ADD_N, 5>(ctx, z, scalars);
        case 6: return launch_zip<T, NDB_OP_ADD_N, 6>(ctx, z, scalars);
        case 7: return launch_zip<T, NDB_OP_ADD_N, 7>(ctx, z, scalars);
        case 8: return launch_zip<T, NDB_OP_ADD_N, 8>(ctx, z, scalars);
        default: break;
        }
        break;
    default: break;
    }
    set_error("Zip: op %d with %d inputs has no kernel", z.op, nin);
    return NDB_ERR_UNSUPPORTED;
}

int zip_dispatch_f32(ndb_ctx *ctx, const ZipPlan &z, const void *scalars);
int zip_dispatch_f64(ndb_ctx *ctx, const ZipPlan &z, const void *scalars);
int zip_dispatch_i32(ndb_ctx *ctx, const ZipPlan &z, const void *scalars);
int zip_dispatch_i64(ndb_ctx *ctx, const ZipPlan &z, const void *scalars);

}  // namespace ndb
