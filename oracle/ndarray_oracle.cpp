// ndarray_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of the algorithms of rust-ndarray 0.17.2 on the hot path (Zip, reductions,
// general_mat_mul).  It exists so that the CUDA library can be checked against "what ndarray would
// have computed" on a machine with no Rust toolchain.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load it; nothing under ndarray_b200/ does.
//
// Parity pinning: the restatement is checked against the literal expected values of ndarray's own
// tests (tests/numeric.rs, tests/azip.rs, tests/array.rs, tests/oper.rs, src/slice.rs doc-tests; see
// tests/test_oracle_golden.py).  GEMM: the arithmetic of `matrixmultiply` 0.3.10 (Cargo.lock:414) is
// not vendored in the reference tree, so GEMM is pinned to tolerance only, against the same naive and
// Kahan references ndarray's tests use (tests/oper.rs:291, crates/numeric-tests/tests/accuracy.rs:25-66).
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).  -ffp-contract=off
// matters: Rust never fuses a*b+c.
#include <math.h>
#include <cmath>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <type_traits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_NDIM 8
#define ORC_MAX_PARTS 9

extern "C" {
typedef struct {
    void *ptr;  // HOST pointer to the logical first element
    int32_t ndim;
    int32_t dtype;  // 0 f32, 1 f64, 2 i32, 3 i64
    int64_t dim[ORC_MAX_NDIM];
    int64_t stride[ORC_MAX_NDIM];  // signed, in elements
} orc_view;
}

enum { ORC_OK = 0, ORC_SHAPE = 1, ORC_OOB = 4, ORC_UNSUPPORTED = 5, ORC_INVALID = 7, ORC_EMPTY = 8, ORC_DIVZERO = 9 };
enum { F32 = 0, F64 = 1, I32 = 2, I64 = 3 };
// op / redop numbering is shared with include/ndb200.h on purpose (the tests pass the same integers)
enum {
    OP_COPY, OP_FILL, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_REM, OP_MUL_ADD, OP_ADD_S, OP_SUB_S, OP_MUL_S, OP_DIV_S,
    OP_REM_S, OP_RSUB_S, OP_RDIV_S, OP_NEG, OP_AXPY, OP_ADD_N, OP_ABS, OP_SQRT, OP_MIN, OP_MAX, OP_SCALE_ADD,
    OP_TF32_HI, OP_TF32_LO, OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE, OP_MATH, OP_CLAMP
};
enum {
    FN_FLOOR, FN_CEIL, FN_ROUND, FN_TRUNC, FN_FRACT, FN_SIGNUM, FN_RECIP, FN_EXP, FN_EXP2, FN_EXP_M1, FN_LN, FN_LOG2, FN_LOG10,
    FN_LN_1P, FN_CBRT, FN_SIN, FN_COS, FN_TAN, FN_ASIN, FN_ACOS, FN_ATAN, FN_SINH, FN_COSH, FN_TANH, FN_ASINH, FN_ACOSH, FN_ATANH,
    FN_TO_DEGREES, FN_TO_RADIANS, FN_POW2, FN_POWI, FN_POWF, FN_LOG, FN_ABS_SUB, FN_HYPOT, FN_IS_NAN, FN_IS_INFINITE
};
enum { RED_SUM, RED_PROD, RED_MIN, RED_MAX, RED_MEAN, RED_VAR, RED_STD };

static inline int64_t esize(int dt) { return (dt == F32 || dt == I32) ? 4 : 8; }
static inline int64_t vsize(const orc_view *v)
{
    int64_t n = 1;
    for (int i = 0; i < v->ndim; ++i) n *= v->dim[i];
    return n;
}

// ------------------------------------------------------------------------------------------------
// shape / stride algebra

// src/dimension/mod.rs:389-397 abs_index, :399-430 to_abs_slice, :445-495 do_slice
extern "C" int orc_slice_axis(orc_view *v, int axis, int64_t start, int64_t end, int has_end, int64_t step)
{
    if (axis < 0 || axis >= v->ndim) return ORC_OOB;
    const int64_t len = v->dim[axis];
    int64_t s = start < 0 ? len + start : start;
    int64_t e0 = has_end ? end : len;
    int64_t e = e0 < 0 ? len + e0 : e0;
    if (e < s) e = s;
    if (s < 0 || s > len || e < 0 || e > len) return ORC_OOB;  // ndassert panics
    if (step == 0) return ORC_INVALID;
    const int64_t m = e - s;
    const int64_t st = v->stride[axis];
    int64_t offset;
    if (m == 0) offset = 0;
    else if (step < 0) offset = (e - 1) * st;
    else offset = s * st;
    const int64_t abs_step = step < 0 ? -step : step;
    int64_t nd;
    if (abs_step == 1) nd = m;
    else {
        nd = m / abs_step;
        if (m % abs_step > 0) nd += 1;
    }
    v->dim[axis] = nd;
    v->stride[axis] = nd <= 1 ? 0 : st * step;
    v->ptr = (char *)v->ptr + offset * esize(v->dtype);
    return ORC_OK;
}

// src/impl_methods.rs:2382-2424 upcast (inside broadcast)
extern "C" int orc_broadcast(const orc_view *v, int ndim, const int64_t *dim, orc_view *out)
{
    if (ndim < v->ndim || ndim > ORC_MAX_NDIM) return ORC_SHAPE;
    orc_view r;
    memset(&r, 0, sizeof(r));
    r.ptr = v->ptr;
    r.dtype = v->dtype;
    r.ndim = ndim;
    for (int i = 0; i < ndim; ++i) r.dim[i] = dim[i];
    int di = ndim - 1;
    for (int fi = v->ndim - 1; fi >= 0; --fi, --di) {  // walk both from the back
        if (dim[di] == v->dim[fi]) r.stride[di] = v->stride[fi];
        else if (v->dim[fi] == 1) r.stride[di] = 0;
        else return ORC_SHAPE;
    }
    for (; di >= 0; --di) r.stride[di] = 0;
    *out = r;
    return ORC_OK;
}

// src/impl_methods.rs:2676-2690 invert_axis
extern "C" int orc_invert_axis(orc_view *v, int axis)
{
    if (axis < 0 || axis >= v->ndim) return ORC_OOB;
    const int64_t m = v->dim[axis];
    if (m != 0) {
        v->ptr = (char *)v->ptr + (m - 1) * v->stride[axis] * esize(v->dtype);
        v->stride[axis] = -v->stride[axis];
    }
    return ORC_OK;
}

// src/dimension/mod.rs:687-711
static bool is_layout_c(const orc_view *v)
{
    if (v->ndim == 1) return v->stride[0] == 1 || v->dim[0] <= 1;
    for (int i = 0; i < v->ndim; ++i)
        if (v->dim[i] == 0) return true;
    int64_t contig = 1;
    for (int i = v->ndim - 1; i >= 0; --i)
        if (v->dim[i] != 1) {
            if (v->stride[i] != contig) return false;
            contig *= v->dim[i];
        }
    return true;
}
// src/dimension/mod.rs:713-737
static bool is_layout_f(const orc_view *v)
{
    if (v->ndim == 1) return v->stride[0] == 1 || v->dim[0] <= 1;
    for (int i = 0; i < v->ndim; ++i)
        if (v->dim[i] == 0) return true;
    int64_t contig = 1;
    for (int i = 0; i < v->ndim; ++i)
        if (v->dim[i] != 1) {
            if (v->stride[i] != contig) return false;
            contig *= v->dim[i];
        }
    return true;
}

enum { CORDER = 1, FORDER = 2, CPREFER = 4, FPREFER = 8 };  // src/layout/mod.rs:13-16

// src/zip/mod.rs:54-77 array_layout
extern "C" uint32_t orc_layout(const orc_view *v)
{
    const int n = v->ndim;
    if (is_layout_c(v)) {
        int big = 0;
        for (int i = 0; i < n; ++i) big += v->dim[i] > 1;
        if (n <= 1 || big <= 1) return CORDER | CPREFER | FORDER | FPREFER;
        return CORDER | CPREFER;
    } else if (n > 1 && is_layout_f(v)) {
        return FORDER | FPREFER;
    } else if (n > 1) {
        if (v->dim[0] > 1 && v->stride[0] == 1) return FPREFER;
        if (v->dim[n - 1] > 1 && v->stride[n - 1] == 1) return CPREFER;
        return 0;
    }
    return 0;
}
// src/layout/mod.rs:76-81
static int tendency(uint32_t l)
{
    return (int)!!(l & CORDER) - (int)!!(l & FORDER) + (int)!!(l & CPREFER) - (int)!!(l & FPREFER);
}

// src/dimension/dimension_trait.rs:294-323 is_contiguous (default_strides :128-150)
extern "C" int orc_is_contiguous(const orc_view *v)
{
    const int n = v->ndim;
    bool nonempty = true;
    for (int i = 0; i < n; ++i) nonempty &= v->dim[i] != 0;
    int64_t def[ORC_MAX_NDIM];
    int64_t c = 1;
    for (int i = n - 1; i >= 0; --i) {
        def[i] = nonempty ? c : 0;
        c *= v->dim[i];
    }
    bool same = true;
    for (int i = 0; i < n; ++i) same &= def[i] == v->stride[i];
    if (same) return 1;
    if (n == 1) return v->dim[0] <= 1 || v->stride[0] == -1;
    int order[ORC_MAX_NDIM];
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order, order + n, [&](int a, int b) { return llabs(v->stride[a]) < llabs(v->stride[b]); });
    int64_t cs = 1;
    for (int k = 0; k < n; ++k) {
        const int i = order[k];
        if (v->dim[i] != 1 && llabs(v->stride[i]) != cs) return 0;
        cs *= v->dim[i];
    }
    return 1;
}

// src/dimension/dimension_trait.rs:343-357 (generic) and :715-724 (Ix2)
extern "C" int orc_min_stride_axis(const orc_view *v)
{
    const int n = v->ndim;
    if (n == 0) return -1;
    if (n == 1) return 0;
    if (n == 2) return llabs(v->stride[0]) < llabs(v->stride[1]) ? 0 : 1;
    int best = -1;  // axes_of(..).rev().min_by_key: first minimum when walking from the LAST axis
    for (int i = n - 1; i >= 0; --i)
        if (best < 0 || llabs(v->stride[i]) < llabs(v->stride[best])) best = i;
    return best;
}

// src/dimension/dimension_trait.rs:359-372: max |stride| among axes of length > 1; LAST maximum wins
// (Iterator::max_by_key returns the last of equal maxima); default Axis(0).
extern "C" int orc_max_stride_axis(const orc_view *v)
{
    const int n = v->ndim;
    if (n == 0) return -1;
    if (n == 1) return 0;
    int best = -1;
    for (int i = 0; i < n; ++i) {
        if (v->dim[i] <= 1) continue;
        if (best < 0 || llabs(v->stride[i]) >= llabs(v->stride[best])) best = i;
    }
    return best < 0 ? 0 : best;
}

// src/dimension/mod.rs:431-443
static int64_t offset_from_low_addr(const orc_view *v)
{
    int64_t off = 0;
    for (int i = 0; i < v->ndim; ++i)
        if (v->stride[i] < 0 && v->dim[i] > 1) off -= v->stride[i] * (v->dim[i] - 1);
    return off;
}

// Zip state as accumulated by Zip::from + build_and (src/zip/mod.rs:212-225,747-758)
struct ZipState {
    uint32_t layout;
    int tend;
};
static ZipState zip_state(int n, const orc_view *parts)
{
    ZipState z{0xF, 0};
    for (int i = 0; i < n; ++i) {
        const uint32_t l = orc_layout(&parts[i]);
        z.layout &= l;
        z.tend += tendency(l);
    }
    return z;
}
// src/zip/mod.rs:278-281 prefer_f
static bool prefer_f(const ZipState &z) { return !(z.layout & CORDER) && ((z.layout & FORDER) || z.tend < 0); }

extern "C" int orc_zip_prefer_f(int n, const orc_view *inputs) { return prefer_f(zip_state(n, inputs)) ? 1 : 0; }

// ------------------------------------------------------------------------------------------------
// elementwise arithmetic with Rust semantics: wrapping integers (release profile), IEEE floats,
// no fusion; integer / and % "panic" (status) on zero divisor and MIN / -1.
template <typename T>
struct Un { using type = T; };
template <>
struct Un<int32_t> { using type = uint32_t; };
template <>
struct Un<int64_t> { using type = uint64_t; };
template <typename T>
static inline T wadd(T a, T b) { using U = typename Un<T>::type; return (T)((U)a + (U)b); }
template <typename T>
static inline T wsub(T a, T b) { using U = typename Un<T>::type; return (T)((U)a - (U)b); }
template <typename T>
static inline T wmul(T a, T b) { using U = typename Un<T>::type; return (T)((U)a * (U)b); }
template <typename T>
static inline T wneg(T a)
{
    if constexpr (std::is_floating_point<T>::value) return -a;
    else { using U = typename Un<T>::type; return (T)((U)0 - (U)a); }
}
template <typename T>
static inline bool int_div_traps(T a, T b)
{
    if constexpr (std::is_floating_point<T>::value) return false;
    else {
        using U = typename Un<T>::type;
        return b == 0 || (b == (T)-1 && a == (T)((U)1 << (sizeof(T) * 8 - 1)));
    }
}
template <typename T>
static inline T wdiv(T a, T b, int *trap)
{
    if (int_div_traps(a, b)) { *trap = 1; return 0; }
    return a / b;
}
template <typename T>
static inline T wrem(T a, T b, int *trap)
{
    if constexpr (std::is_same<T, float>::value) return fmodf(a, b);
    else if constexpr (std::is_same<T, double>::value) return fmod(a, b);
    else {
        if (int_div_traps(a, b)) { *trap = 1; return 0; }
        return a % b;
    }
}
// f32::min / f32::max ignore a NaN operand; the sign of a zero result is unspecified in Rust, so the oracle and
// the kernels share one explicit rule: the first operand wins ties.
template <typename T>
static inline T wmin(T a, T b)
{
    if constexpr (std::is_floating_point<T>::value) return (b < a || a != a) ? b : a;
    else return b < a ? b : a;
}
template <typename T>
static inline T wmax(T a, T b)
{
    if constexpr (std::is_floating_point<T>::value) return (b > a || a != a) ? b : a;
    else return b > a ? b : a;
}

// the float maps of src/numeric/impl_float_maths.rs:80-160 (each is `mapv(A::f)` on num_traits::Float, i.e. libm)
template <typename T>
static inline T math_fn(int fn, T x, T p)
{
    if constexpr (std::is_floating_point<T>::value) {
        switch (fn) {
        case FN_FLOOR: return std::floor(x);
        case FN_CEIL: return std::ceil(x);
        case FN_ROUND: return std::round(x);
        case FN_TRUNC: return std::trunc(x);
        case FN_FRACT: return x - std::trunc(x);
        case FN_SIGNUM: return x != x ? x : (T)copysign(1.0, (double)x);
        case FN_RECIP: return (T)1 / x;
        case FN_EXP: return std::exp(x);
        case FN_EXP2: return std::exp2(x);
        case FN_EXP_M1: return std::expm1(x);
        case FN_LN: return std::log(x);
        case FN_LOG2: return std::log2(x);
        case FN_LOG10: return std::log10(x);
        case FN_LN_1P: return std::log1p(x);
        case FN_CBRT: return std::cbrt(x);
        case FN_SIN: return std::sin(x);
        case FN_COS: return std::cos(x);
        case FN_TAN: return std::tan(x);
        case FN_ASIN: return std::asin(x);
        case FN_ACOS: return std::acos(x);
        case FN_ATAN: return std::atan(x);
        case FN_SINH: return std::sinh(x);
        case FN_COSH: return std::cosh(x);
        case FN_TANH: return std::tanh(x);
        case FN_ASINH: return std::asinh(x);
        case FN_ACOSH: return std::acosh(x);
        case FN_ATANH: return std::atanh(x);
        case FN_TO_DEGREES:
            return std::is_same<T, float>::value ? (T)((float)x * 57.2957795130823208767981548141051703f)
                                                 : (T)((double)x * (180.0 / 3.14159265358979323846264338327950288));
        case FN_TO_RADIANS:
            return std::is_same<T, float>::value ? (T)((float)x * (3.14159265358979323846264338327950288f / 180.0f))
                                                 : (T)((double)x * (3.14159265358979323846264338327950288 / 180.0));
        case FN_POW2: return x * x;
        case FN_POWI: {  // f32::powi -> llvm.powi -> compiler-rt __powisf2 / __powidf2
            int b = (int)p;
            const bool recip = b < 0;
            T a = x, r = (T)1;
            for (;;) {
                if (b & 1) r *= a;
                b /= 2;
                if (b == 0) break;
                a *= a;
            }
            return recip ? (T)1 / r : r;
        }
        case FN_POWF: return std::pow(x, p);
        case FN_LOG: return std::log(x) / std::log(p);
        case FN_ABS_SUB: return std::fdim(x, p);
        case FN_HYPOT: return std::hypot(x, p);
        case FN_IS_NAN: return x != x ? (T)1 : (T)0;
        case FN_IS_INFINITE: return std::isinf(x) ? (T)1 : (T)0;
        }
    }
    return x;
}

template <typename T>
__attribute__((always_inline)) static inline T apply(int op, int nin, const T *x, T s0, T s1, int *trap)
{
    switch (op) {
    case OP_COPY: return x[0];
    case OP_FILL: return s0;
    case OP_ADD: return wadd(x[0], x[1]);
    case OP_SUB: return wsub(x[0], x[1]);
    case OP_MUL: return wmul(x[0], x[1]);
    case OP_DIV: return wdiv(x[0], x[1], trap);
    case OP_REM: return wrem(x[0], x[1], trap);
    case OP_MUL_ADD: return wadd(wmul(x[0], x[1]), x[2]);
    case OP_ADD_S: return wadd(x[0], s0);
    case OP_SUB_S: return wsub(x[0], s0);
    case OP_MUL_S: return wmul(x[0], s0);
    case OP_DIV_S: return wdiv(x[0], s0, trap);
    case OP_REM_S: return wrem(x[0], s0, trap);
    case OP_RSUB_S: return wsub(s0, x[0]);
    case OP_RDIV_S: return wdiv(s0, x[0], trap);
    case OP_NEG: return wneg(x[0]);
    case OP_AXPY: return wadd(x[0], wmul(s0, x[1]));  // scaled_add: *a = *a + alpha * b  (impl_linalg.rs:404)
    case OP_ADD_N: {
        T acc = wadd(x[0], x[1]);
        for (int i = 2; i < nin; ++i) acc = wadd(acc, x[i]);
        return acc;
    }
    case OP_ABS:
        if constexpr (std::is_floating_point<T>::value) return (T)fabs((double)x[0]);
        else return x[0] < 0 ? wneg(x[0]) : x[0];
    case OP_SQRT:
        if constexpr (std::is_same<T, float>::value) return sqrtf(x[0]);
        else if constexpr (std::is_same<T, double>::value) return sqrt(x[0]);
        else return x[0];
    case OP_MIN: return wmin(x[0], x[1]);
    case OP_MAX: return wmax(x[0], x[1]);
    case OP_SCALE_ADD: return wadd(wmul(x[0], s0), wmul(x[1], s1));
    case OP_EQ: return x[0] == x[1] ? (T)1 : (T)0;
    case OP_NE: return x[0] != x[1] ? (T)1 : (T)0;
    case OP_LT: return x[0] < x[1] ? (T)1 : (T)0;
    case OP_LE: return x[0] <= x[1] ? (T)1 : (T)0;
    case OP_GT: return x[0] > x[1] ? (T)1 : (T)0;
    case OP_GE: return x[0] >= x[1] ? (T)1 : (T)0;
    case OP_MATH: return math_fn<T>((int)s0, x[0], s1);
    case OP_CLAMP: return x[0] != x[0] ? x[0] : (x[0] < s0 ? s0 : (x[0] > s1 ? s1 : x[0]));
    default: return x[0];
    }
}

// The traversal of Zip::for_each_core (src/zip/mod.rs:309-424) for one Zip over `n_parts` producers
// of identical shape.  parts[0] is written, the rest are read.
// OP and NIN are compile-time so that, like the monomorphised Rust closure, the element function inlines and the
// contiguous loops vectorise (NIN < 0: the n-ary sum, arity known at run time).
template <typename T, int OP, int NIN>
static void zip_serial_ct(const void *scalars, int n_parts, const orc_view *parts, int *trap)
{
    constexpr int op = OP;
    const int nin = NIN >= 0 ? NIN : n_parts - 1;
    T s0 = 0, s1 = 0;
    if (scalars) {
        s0 = ((const T *)scalars)[0];
        if (op == OP_SCALE_ADD || op == OP_MATH || op == OP_CLAMP) s1 = ((const T *)scalars)[1];
    }
    const orc_view &o = parts[0];
    const int nd = o.ndim;
    const int64_t size = vsize(&o);
    if (size == 0) return;
    T *ptr[ORC_MAX_PARTS];
    for (int p = 0; p < n_parts; ++p) ptr[p] = (T *)parts[p].ptr;
    T x[ORC_MAX_PARTS];
    (void)x;
    if (nd == 0) {  // :313-316
        for (int p = 0; p < nin; ++p) x[p] = *ptr[p + 1];
        *ptr[0] = apply<T>(op, nin, x, s0, s1, trap);
        return;
    }
    const ZipState zs = zip_state(n_parts, parts);
    if (zs.layout & (CORDER | FORDER)) {
        // for_each_core_contiguous :323-341 -> inner :343-357 with unit strides
        if constexpr (NIN >= 0 && NIN <= 3) {
            T *__restrict__ o_ = ptr[0];
            const T *a_ = NIN > 0 ? ptr[1] : nullptr, *b_ = NIN > 1 ? ptr[2] : nullptr, *c_ = NIN > 2 ? ptr[3] : nullptr;
            const bool alias = (NIN > 0 && a_ == o_) || (NIN > 1 && b_ == o_) || (NIN > 2 && c_ == o_);
            if (!alias) {
#pragma omp simd
                for (int64_t i = 0; i < size; ++i) {
                    T v[3] = {NIN > 0 ? a_[i] : T(0), NIN > 1 ? b_[i] : T(0), NIN > 2 ? c_[i] : T(0)};
                    o_[i] = apply<T>(op, nin, v, s0, s1, trap);
                }
                return;
            }
            T *o2 = ptr[0];  // in-place form (a = b*c + a): same element read and written, still independent across i
#pragma omp simd
            for (int64_t i = 0; i < size; ++i) {
                T v[3] = {NIN > 0 ? ptr[1][i] : T(0), NIN > 1 ? ptr[2][i] : T(0), NIN > 2 ? ptr[3][i] : T(0)};
                o2[i] = apply<T>(op, nin, v, s0, s1, trap);
            }
            return;
        }
        for (int64_t i = 0; i < size; ++i) {
            for (int p = 0; p < nin; ++p) x[p] = ptr[p + 1][i];
            ptr[0][i] = apply<T>(op, nin, x, s0, s1, trap);
        }
        return;
    }
    // for_each_core_strided :359-373: unroll the last axis (C) or axis 0 (F, when tendency < 0)
    const bool f_order = nd > 1 && zs.tend < 0;
    const int unroll_axis = f_order ? 0 : nd - 1;
    const int64_t inner_len = o.dim[unroll_axis];
    int64_t inner_stride[ORC_MAX_PARTS];
    for (int p = 0; p < n_parts; ++p) inner_stride[p] = parts[p].stride[unroll_axis];
    int64_t idx[ORC_MAX_NDIM] = {0};
    // odometer over the remaining axes: next_for (C: last axis fastest, dimension_trait.rs:200-222)
    // or next_for_f (F: first axis fastest, :225-240)
    for (;;) {
        int64_t off[ORC_MAX_PARTS];
        for (int p = 0; p < n_parts; ++p) {
            off[p] = 0;
            for (int k = 0; k < nd; ++k) off[p] += idx[k] * parts[p].stride[k];
        }
        if constexpr (NIN >= 0 && NIN <= 3) {
            T *o_ = ptr[0] + off[0];
            const int64_t so = inner_stride[0];
            const T *a_ = NIN > 0 ? ptr[1] + off[1] : nullptr, *b_ = NIN > 1 ? ptr[2] + off[2] : nullptr,
                    *c_ = NIN > 2 ? ptr[3] + off[3] : nullptr;
            const int64_t sa = NIN > 0 ? inner_stride[1] : 0, sb = NIN > 1 ? inner_stride[2] : 0, sc = NIN > 2 ? inner_stride[3] : 0;
            for (int64_t i = 0; i < inner_len; ++i) {
                T v[3] = {NIN > 0 ? a_[i * sa] : T(0), NIN > 1 ? b_[i * sb] : T(0), NIN > 2 ? c_[i * sc] : T(0)};
                o_[i * so] = apply<T>(op, nin, v, s0, s1, trap);
            }
        } else {
            for (int64_t i = 0; i < inner_len; ++i) {
                for (int p = 0; p < nin; ++p) x[p] = ptr[p + 1][off[p + 1] + i * inner_stride[p + 1]];
                ptr[0][off[0] + i * inner_stride[0]] = apply<T>(op, nin, x, s0, s1, trap);
            }
        }
        bool done = true;
        if (!f_order) {
            for (int k = nd - 2; k >= 0; --k) {
                if (++idx[k] < o.dim[k]) { done = false; break; }
                idx[k] = 0;
            }
        } else {
            for (int k = 1; k < nd; ++k) {
                if (++idx[k] < o.dim[k]) { done = false; break; }
                idx[k] = 0;
            }
        }
        if (done) break;
    }
}

template <typename T>
static void zip_serial(int op, const void *scalars, int n_parts, const orc_view *parts, int *trap)
{
    switch (op) {
#define ORC_CASE(OP, NIN) \
    case OP: zip_serial_ct<T, OP, NIN>(scalars, n_parts, parts, trap); return;
        ORC_CASE(OP_COPY, 1) ORC_CASE(OP_FILL, 0) ORC_CASE(OP_ADD, 2) ORC_CASE(OP_SUB, 2) ORC_CASE(OP_MUL, 2)
        ORC_CASE(OP_DIV, 2) ORC_CASE(OP_REM, 2) ORC_CASE(OP_MUL_ADD, 3) ORC_CASE(OP_ADD_S, 1) ORC_CASE(OP_SUB_S, 1)
        ORC_CASE(OP_MUL_S, 1) ORC_CASE(OP_DIV_S, 1) ORC_CASE(OP_REM_S, 1) ORC_CASE(OP_RSUB_S, 1) ORC_CASE(OP_RDIV_S, 1)
        ORC_CASE(OP_NEG, 1) ORC_CASE(OP_AXPY, 2) ORC_CASE(OP_ADD_N, -1) ORC_CASE(OP_ABS, 1) ORC_CASE(OP_SQRT, 1)
        ORC_CASE(OP_MIN, 2) ORC_CASE(OP_MAX, 2) ORC_CASE(OP_SCALE_ADD, 2) ORC_CASE(OP_EQ, 2) ORC_CASE(OP_NE, 2)
        ORC_CASE(OP_LT, 2) ORC_CASE(OP_LE, 2) ORC_CASE(OP_GT, 2) ORC_CASE(OP_GE, 2) ORC_CASE(OP_MATH, 1) ORC_CASE(OP_CLAMP, 1)
#undef ORC_CASE
    default: break;
    }
}

// Zip::split (src/zip/mod.rs:811-815,883-895): halve along max_stride_axis (first len>1 axis for C,
// last for F).  rayon's par_for_each (src/parallel/par.rs:270-297) applies it recursively; a static
// split into `threads` slabs along that axis is the same partition with the stealing removed.
template <typename T>
static void zip_typed(int op, const void *scalars, int n_parts, const orc_view *parts, int threads, int *trap)
{
    const orc_view &o = parts[0];
    if (threads <= 1 || o.ndim == 0 || vsize(&o) < 65536) {
        zip_serial<T>(op, scalars, n_parts, parts, trap);
        return;
    }
    const ZipState zs = zip_state(n_parts, parts);
    int axis = 0;
    if (prefer_f(zs)) {
        axis = o.ndim - 1;
        for (int i = o.ndim - 1; i >= 0; --i)
            if (o.dim[i] > 1) { axis = i; break; }
    } else {
        axis = 0;
        for (int i = 0; i < o.ndim; ++i)
            if (o.dim[i] > 1) { axis = i; break; }
    }
    const int64_t len = o.dim[axis];
    const int nt = (int)std::min<int64_t>(threads, len);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; ++t) {
        const int64_t lo = len * t / nt, hi = len * (t + 1) / nt;
        orc_view sub[ORC_MAX_PARTS];
        for (int p = 0; p < n_parts; ++p) {
            sub[p] = parts[p];
            sub[p].dim[axis] = hi - lo;
            sub[p].ptr = (char *)parts[p].ptr + lo * parts[p].stride[axis] * (int64_t)sizeof(T);
        }
        int local_trap = 0;
        zip_serial<T>(op, scalars, n_parts, sub, &local_trap);
        if (local_trap) {
#pragma omp atomic write
            *trap = 1;
        }
    }
}

extern "C" int orc_zip(int op, const void *scalars, int n_parts, const orc_view *parts, int threads)
{
    if (n_parts < 1 || n_parts > ORC_MAX_PARTS) return ORC_UNSUPPORTED;
    for (int p = 1; p < n_parts; ++p) {  // zip_dimension_check, src/zip/mod.rs:247-258
        if (parts[p].ndim != parts[0].ndim) return ORC_SHAPE;
        for (int i = 0; i < parts[0].ndim; ++i)
            if (parts[p].dim[i] != parts[0].dim[i]) return ORC_SHAPE;
    }
    int trap = 0;
    switch (parts[0].dtype) {
    case F32: zip_typed<float>(op, scalars, n_parts, parts, threads, &trap); break;
    case F64: zip_typed<double>(op, scalars, n_parts, parts, threads, &trap); break;
    case I32: zip_typed<int32_t>(op, scalars, n_parts, parts, threads, &trap); break;
    case I64: zip_typed<int64_t>(op, scalars, n_parts, parts, threads, &trap); break;
    default: return ORC_UNSUPPORTED;
    }
    return trap ? ORC_DIVZERO : ORC_OK;
}

// ------------------------------------------------------------------------------------------------
// reductions

// src/numeric_util.rs:14-51 unrolled_fold
template <typename T, typename F>
static T unrolled_fold(const T *xs, int64_t len, T init, F f)
{
    T acc = init;
    T p0 = init, p1 = init, p2 = init, p3 = init, p4 = init, p5 = init, p6 = init, p7 = init;
    while (len >= 8) {
        p0 = f(p0, xs[0]);
        p1 = f(p1, xs[1]);
        p2 = f(p2, xs[2]);
        p3 = f(p3, xs[3]);
        p4 = f(p4, xs[4]);
        p5 = f(p5, xs[5]);
        p6 = f(p6, xs[6]);
        p7 = f(p7, xs[7]);
        xs += 8;
        len -= 8;
    }
    acc = f(acc, f(p0, p4));
    acc = f(acc, f(p1, p5));
    acc = f(acc, f(p2, p6));
    acc = f(acc, f(p3, p7));
    for (int64_t i = 0; i < len && i < 7; ++i) acc = f(acc, xs[i]);
    return acc;
}

// src/numeric_util.rs:56-93 unrolled_dot
template <typename T>
static T unrolled_dot(const T *xs, const T *ys, int64_t len)
{
    T sum = 0;
    T p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (len >= 8) {
        for (int j = 0; j < 8; ++j) p[j] = wadd(p[j], wmul(xs[j], ys[j]));
        xs += 8;
        ys += 8;
        len -= 8;
    }
    sum = wadd(sum, wadd(p[0], p[4]));
    sum = wadd(sum, wadd(p[1], p[5]));
    sum = wadd(sum, wadd(p[2], p[6]));
    sum = wadd(sum, wadd(p[3], p[7]));
    for (int64_t i = 0; i < len && i < 7; ++i) sum = wadd(sum, wmul(xs[i], ys[i]));
    return sum;
}

// ArrayRef::sum / product (src/numeric/impl_numeric.rs:32-47, 83-98): one unrolled_fold over the whole
// block in MEMORY order when contiguous (as_slice_memory_order, src/impl_methods.rs:1843-1851), else
// per-row (last-axis lanes, logical order) partials added sequentially.
template <typename T, typename F>
static T fold_all(const orc_view *v, T init, F f)
{
    const int64_t n = vsize(v);
    if (orc_is_contiguous(v)) {
        const T *low = (const T *)v->ptr - offset_from_low_addr(v);
        return unrolled_fold<T>(low, n, init, f);
    }
    T total = init;
    if (n == 0) return total;
    const int nd = v->ndim;  // not contiguous implies nd >= 1
    const int64_t rl = v->dim[nd - 1], rs = v->stride[nd - 1];
    int64_t idx[ORC_MAX_NDIM] = {0};
    for (;;) {
        int64_t off = 0;
        for (int k = 0; k < nd - 1; ++k) off += idx[k] * v->stride[k];
        const T *row = (const T *)v->ptr + off;
        T part;
        if (rl <= 1 || rs == 1) {  // row.as_slice()
            part = unrolled_fold<T>(row, rl, init, f);
        } else {
            part = init;
            for (int64_t i = 0; i < rl; ++i) part = f(part, row[i * rs]);
        }
        total = f(total, part);
        bool done = true;
        for (int k = nd - 2; k >= 0; --k) {
            if (++idx[k] < v->dim[k]) { done = false; break; }
            idx[k] = 0;
        }
        if (done) break;
    }
    return total;
}

template <typename T>
static int reduce_all_typed(int redop, const orc_view *v, void *out)
{
    T r;
    switch (redop) {
    case RED_SUM: r = fold_all<T>(v, (T)0, [](T a, T b) { return wadd(a, b); }); break;
    case RED_PROD: r = fold_all<T>(v, (T)1, [](T a, T b) { return wmul(a, b); }); break;
    default: return ORC_UNSUPPORTED;
    }
    *(T *)out = r;
    return ORC_OK;
}

extern "C" int orc_reduce_all(int redop, const orc_view *v, void *out)
{
    switch (v->dtype) {
    case F32: return reduce_all_typed<float>(redop, v, out);
    case F64: return reduce_all_typed<double>(redop, v, out);
    case I32: return reduce_all_typed<int32_t>(redop, v, out);
    case I64: return reduce_all_typed<int64_t>(redop, v, out);
    }
    return ORC_UNSUPPORTED;
}

static void remove_axis(const orc_view *v, int axis, orc_view *out)
{
    *out = *v;
    for (int i = axis; i + 1 < v->ndim; ++i) {
        out->dim[i] = v->dim[i + 1];
        out->stride[i] = v->stride[i + 1];
    }
    out->ndim = v->ndim - 1;
}

// visit every index of `shape` view in logical (C) order, giving element offsets into two views
template <typename Fn>
static void for_each_index(const orc_view *a, const orc_view *b, Fn fn)
{
    const int nd = a->ndim;
    if (vsize(a) == 0) return;
    if (nd == 1) {  // the common shape of a sum_axis result: one tight loop (what zip_mut_with_same_shape's slice loop is)
        const int64_t n = a->dim[0], sa = a->stride[0], sb = b->stride[0];
        if (sa == 1 && sb == 1) {
            for (int64_t i = 0; i < n; ++i) fn(i, i);
        } else {
            for (int64_t i = 0; i < n; ++i) fn(i * sa, i * sb);
        }
        return;
    }
    int64_t idx[ORC_MAX_NDIM] = {0};
    for (;;) {
        int64_t oa = 0, ob = 0;
        for (int k = 0; k < nd; ++k) {
            oa += idx[k] * a->stride[k];
            ob += idx[k] * b->stride[k];
        }
        fn(oa, ob);
        bool done = true;
        for (int k = nd - 1; k >= 0; --k) {
            if (++idx[k] < a->dim[k]) { done = false; break; }
            idx[k] = 0;
        }
        if (done) break;
    }
}

template <typename T>
static int reduce_axis_typed(int redop, int axis, const orc_view *in, const orc_view *out, double ddof_d)
{
    const int64_t n = in->dim[axis];
    const int64_t s = in->stride[axis];
    orc_view base;  // `in` with `axis` removed: where each lane starts
    remove_axis(in, axis, &base);
    T *optr = (T *)out->ptr;
    const T *iptr = (const T *)in->ptr;

    if (redop == RED_SUM || redop == RED_PROD || redop == RED_MEAN) {
        if (redop == RED_MEAN && n == 0) return ORC_EMPTY;  // mean_axis -> None (impl_numeric.rs:348-350)
        const bool prod = redop == RED_PROD;
        const T init = prod ? (T)1 : (T)0;
        auto f = [prod](T a, T b) { return prod ? wmul(a, b) : wadd(a, b); };
        // sum_axis / product_axis (impl_numeric.rs:268-283, 305-320)
        if (axis == orc_min_stride_axis(in)) {
            // fast path: Zip::from(lanes(axis)).map_collect(|lane| lane.sum()); a lane is a 1-d view, so
            // sum() folds it in memory order when |stride| == 1 (or len <= 1), else element by element.
            for_each_index(&base, out, [&](int64_t ob, int64_t oo) {
                orc_view lane;
                memset(&lane, 0, sizeof(lane));
                lane.ptr = (void *)(iptr + ob);
                lane.ndim = 1;
                lane.dtype = in->dtype;
                lane.dim[0] = n;
                lane.stride[0] = s;
                optr[oo] = fold_all<T>(&lane, init, f);
            });
        } else {
            // slow path: res = zeros; for subview in axis_iter(axis) { res = res + &subview }
            for_each_index(&base, out, [&](int64_t, int64_t oo) { optr[oo] = init; });
            for (int64_t i = 0; i < n; ++i)
                for_each_index(&base, out, [&](int64_t ob, int64_t oo) { optr[oo] = f(optr[oo], iptr[ob + i * s]); });
        }
        if (redop == RED_MEAN) {
            // sum / aview0(&A::from_usize(n)) (impl_numeric.rs:351-354); integer division truncates
            const T nn = (T)n;
            int trap = 0;
            for_each_index(&base, out, [&](int64_t, int64_t oo) { optr[oo] = wdiv(optr[oo], nn, &trap); });
        }
        return ORC_OK;
    }
    if (redop == RED_MIN || redop == RED_MAX) {
        // fold_axis(axis, init, |&a, &b| a.min(b)) (impl_methods.rs:3130-3141): strictly in axis order.
        // init is the identity for the op (max/min value, or +-inf).
        if (n == 0) return ORC_EMPTY;
        for_each_index(&base, out, [&](int64_t ob, int64_t oo) { optr[oo] = iptr[ob]; });
        for (int64_t i = 1; i < n; ++i)
            for_each_index(&base, out, [&](int64_t ob, int64_t oo) {
                const T x = iptr[ob + i * s];
                optr[oo] = redop == RED_MIN ? wmin(optr[oo], x) : wmax(optr[oo], x);
            });
        return ORC_OK;
    }
    if constexpr (std::is_floating_point<T>::value) {
        if (redop == RED_VAR || redop == RED_STD) {
            // var_axis (impl_numeric.rs:401-425): Welford in axis order with a FUSED mul_add.
            const T ddof = (T)ddof_d;
            const T nn = (T)n;
            if (ddof < (T)0 || ddof > nn) return ORC_INVALID;  // assert!(!(ddof < zero || ddof > n)); NaN passes
            const T dof = nn - ddof;
            std::vector<T> mean((size_t)std::max<int64_t>(vsize(out), 1), (T)0), m2(mean.size(), (T)0);
            for (int64_t i = 0; i < n; ++i) {
                const T count = (T)(i + 1);
                size_t j = 0;
                for_each_index(&base, out, [&](int64_t ob, int64_t) {
                    const T x = iptr[ob + i * s];
                    const T delta = x - mean[j];
                    mean[j] = mean[j] + delta / count;
                    if constexpr (std::is_same<T, float>::value) m2[j] = fmaf(x - mean[j], delta, m2[j]);
                    else m2[j] = fma(x - mean[j], delta, m2[j]);
                    ++j;
                });
            }
            size_t j = 0;
            for_each_index(&base, out, [&](int64_t, int64_t oo) {
                T v = m2[j++] / dof;
                if (redop == RED_STD) v = std::is_same<T, float>::value ? (T)sqrtf((float)v) : (T)sqrt((double)v);
                optr[oo] = v;
            });
            return ORC_OK;
        }
    }
    return ORC_UNSUPPORTED;
}

extern "C" int orc_reduce_axis(int redop, int axis, const orc_view *in, const orc_view *out, double ddof)
{
    if (axis < 0 || axis >= in->ndim) return ORC_OOB;
    if (out->ndim != in->ndim - 1) return ORC_SHAPE;
    for (int i = 0, j = 0; i < in->ndim; ++i) {
        if (i == axis) continue;
        if (out->dim[j++] != in->dim[i]) return ORC_SHAPE;
    }
    switch (in->dtype) {
    case F32: return reduce_axis_typed<float>(redop, axis, in, out, ddof);
    case F64: return reduce_axis_typed<double>(redop, axis, in, out, ddof);
    case I32: return reduce_axis_typed<int32_t>(redop, axis, in, out, ddof);
    case I64: return reduce_axis_typed<int64_t>(redop, axis, in, out, ddof);
    }
    return ORC_UNSUPPORTED;
}

// whole-array var (impl_numeric.rs:180-203): Welford over for_each order = logical order for a
// non-contiguous view; for a contiguous one `for_each` walks memory order (impl_methods.rs:3113).
// The golden vectors only use 1-d standard arrays where the two coincide; we take logical order.
extern "C" int orc_var(const orc_view *v, double ddof, double *out)
{
    if (v->dtype != F64) return ORC_UNSUPPORTED;
    const double n = (double)vsize(v);
    if (ddof < 0.0 || ddof > n) return ORC_INVALID;
    const double dof = n - ddof;
    double mean = 0, m2 = 0;
    int64_t i = 0;
    const double *p = (const double *)v->ptr;
    for_each_index(v, v, [&](int64_t oa, int64_t) {
        const double x = p[oa];
        const double count = (double)(i + 1);
        const double delta = x - mean;
        mean = mean + delta / count;
        m2 = fma(x - mean, delta, m2);
        ++i;
    });
    *out = m2 / dof;
    return ORC_OK;
}

// dot_generic (src/linalg/impl_linalg.rs:74-91): unrolled_dot when both are standard slices, else a
// sequential loop.
extern "C" int orc_dot(const orc_view *x, const orc_view *y, void *out)
{
    if (x->ndim != 1 || y->ndim != 1 || x->dim[0] != y->dim[0]) return ORC_SHAPE;
    const int64_t n = x->dim[0];
    const bool slices = (n <= 1 || x->stride[0] == 1) && (n <= 1 || y->stride[0] == 1);
#define ORC_DOT(T)                                                                                   \
    {                                                                                                \
        const T *a = (const T *)x->ptr, *b = (const T *)y->ptr;                                      \
        T r;                                                                                         \
        if (slices) r = unrolled_dot<T>(a, b, n);                                                    \
        else {                                                                                       \
            r = 0;                                                                                   \
            for (int64_t i = 0; i < n; ++i) r = wadd(r, wmul(a[i * x->stride[0]], b[i * y->stride[0]])); \
        }                                                                                            \
        *(T *)out = r;                                                                               \
        return ORC_OK;                                                                               \
    }
    switch (x->dtype) {
    case F32: ORC_DOT(float)
    case F64: ORC_DOT(double)
    case I32: ORC_DOT(int32_t)
    case I64: ORC_DOT(int64_t)
    }
#undef ORC_DOT
    return ORC_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// exact-ish references used only to state tolerances: long-double Neumaier sums.
extern "C" int orc_sum_exact(const orc_view *v, double *out)
{
    long double sum = 0, comp = 0;
    auto add = [&](long double x) {
        const long double t = sum + x;
        if (fabsl(sum) >= fabsl(x)) comp += (sum - t) + x;
        else comp += (x - t) + sum;
        sum = t;
    };
    switch (v->dtype) {
    case F32: for_each_index(v, v, [&](int64_t o, int64_t) { add(((const float *)v->ptr)[o]); }); break;
    case F64: for_each_index(v, v, [&](int64_t o, int64_t) { add(((const double *)v->ptr)[o]); }); break;
    default: return ORC_UNSUPPORTED;
    }
    *out = (double)(sum + comp);
    return ORC_OK;
}

// ------------------------------------------------------------------------------------------------
// general_mat_mul semantics (src/linalg/impl_linalg.rs:633-643): C <- alpha*A*B + beta*C; with beta == 0
// C is not read (:323-329).  The product itself lives in the un-vendored `matrixmultiply` crate, so the
// oracle uses the two references ndarray's own tests use:
//   mode 0: the naive i,j,(0..k).fold of tests/oper.rs:291-318 (reference_mat_mul), combined as
//           alpha * ref + beta * c like gen_mat_mul (:607-611)
//   mode 1: the Kahan-summed reference of crates/numeric-tests/tests/accuracy.rs:25-66
template <typename T>
static int gemm_typed(int mode, T alpha, const orc_view *a, const orc_view *b, T beta, const orc_view *c, int threads)
{
    const int64_t m = a->dim[0], k = a->dim[1], n = b->dim[1];
    const T *ap = (const T *)a->ptr, *bp = (const T *)b->ptr;
    T *cp = (T *)c->ptr;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t i = 0; i < m; ++i) {
        for (int64_t j = 0; j < n; ++j) {
            T acc;
            if (mode == 0) {
                acc = 0;
                for (int64_t x = 0; x < k; ++x)
                    acc = wadd(acc, wmul(ap[i * a->stride[0] + x * a->stride[1]], bp[x * b->stride[0] + j * b->stride[1]]));
            } else {
                T sum = 0, comp = 0;
                for (int64_t x = 0; x < k; ++x) {
                    const T elt = wmul(ap[i * a->stride[0] + x * a->stride[1]], bp[x * b->stride[0] + j * b->stride[1]]);
                    const T y = wsub(elt, comp);
                    const T t = wadd(sum, y);
                    comp = wsub(wsub(t, sum), y);
                    sum = t;
                }
                acc = sum;
            }
            T *dst = cp + i * c->stride[0] + j * c->stride[1];
            if (beta == (T)0) *dst = wmul(alpha, acc);
            else *dst = wadd(wmul(alpha, acc), wmul(beta, *dst));
        }
    }
    return ORC_OK;
}

extern "C" int orc_gemm(int mode, const void *alpha, const orc_view *a, const orc_view *b, const void *beta,
                        const orc_view *c, int threads)
{
    if (a->ndim != 2 || b->ndim != 2 || c->ndim != 2) return ORC_SHAPE;
    if (a->dim[1] != b->dim[0] || c->dim[0] != a->dim[0] || c->dim[1] != b->dim[1]) return ORC_SHAPE;  // :351-355
    switch (a->dtype) {
    case F32: return gemm_typed<float>(mode, *(const float *)alpha, a, b, *(const float *)beta, c, threads);
    case F64: return gemm_typed<double>(mode, *(const double *)alpha, a, b, *(const double *)beta, c, threads);
    case I32: return gemm_typed<int32_t>(mode, *(const int32_t *)alpha, a, b, *(const int32_t *)beta, c, threads);
    case I64: return gemm_typed<int64_t>(mode, *(const int64_t *)alpha, a, b, *(const int64_t *)beta, c, threads);
    }
    return ORC_UNSUPPORTED;
}

// `x as B` with Rust semantics (saturating float -> int, NaN -> 0; wrapping int -> int)
template <typename TO, typename TI>
static inline TO cast_one(TI x)
{
    if constexpr (std::is_floating_point<TI>::value && std::is_integral<TO>::value) {
        if (x != x) return 0;
        const double lo = (double)std::numeric_limits<TO>::min(), hi = (double)std::numeric_limits<TO>::max();
        const double d = std::trunc((double)x);
        if (d <= lo) return std::numeric_limits<TO>::min();
        if (d >= hi) return std::numeric_limits<TO>::max();
        return (TO)d;
    } else if constexpr (std::is_integral<TI>::value && std::is_integral<TO>::value) {
        return (TO)x;  // sign-extend when widening, keep the low bits when narrowing (two's complement, like `as`)
    } else {
        return (TO)x;
    }
}

template <typename TO, typename TI>
static void cast_typed(const orc_view *out, const orc_view *in)
{
    TO *o = (TO *)out->ptr;
    const TI *i = (const TI *)in->ptr;
    for_each_index(out, in, [&](int64_t oo, int64_t oi) { o[oo] = cast_one<TO, TI>(i[oi]); });
}

extern "C" int orc_cast(const orc_view *out, const orc_view *in)
{
    if (out->ndim != in->ndim) return ORC_SHAPE;
    for (int k = 0; k < out->ndim; ++k)
        if (out->dim[k] != in->dim[k]) return ORC_SHAPE;
#define ORC_CAST_IN(TO)                                                  \
    switch (in->dtype) {                                                 \
    case F32: cast_typed<TO, float>(out, in); return ORC_OK;             \
    case F64: cast_typed<TO, double>(out, in); return ORC_OK;            \
    case I32: cast_typed<TO, int32_t>(out, in); return ORC_OK;           \
    case I64: cast_typed<TO, int64_t>(out, in); return ORC_OK;           \
    }                                                                    \
    return ORC_UNSUPPORTED;
    switch (out->dtype) {
    case F32: ORC_CAST_IN(float)
    case F64: ORC_CAST_IN(double)
    case I32: ORC_CAST_IN(int32_t)
    case I64: ORC_CAST_IN(int64_t)
    }
#undef ORC_CAST_IN
    return ORC_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// Synthetic data + BASELINE config 5 (2^32-element f32 `a = b*c + a`, then `a.sum()`).
//
// The counter generator is the harness's, not ndarray's: value(seed, i) is a pure function of the flat index so that any
// shard of a logical array can be regenerated anywhere (same definition as ndarray_b200/csrc/rng.cu, restated here).
static inline uint64_t counter_hash(uint64_t seed, uint64_t i)
{
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float counter_f32(uint64_t seed, uint64_t i) { return (float)(uint32_t)(counter_hash(seed, i) >> 40) * 0x1p-23f - 1.0f; }
static inline double counter_f64(uint64_t seed, uint64_t i) { return (double)(counter_hash(seed, i) >> 11) * 0x1p-52 - 1.0; }

extern "C" int orc_counter_fill(void *dst, int dtype, int64_t n, uint64_t seed, uint64_t first, int threads)
{
    const int nt = threads > 0 ? threads : 1;
    switch (dtype) {
    case F32:
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t i = 0; i < n; ++i) ((float *)dst)[i] = counter_f32(seed, first + (uint64_t)i);
        return ORC_OK;
    case F64:
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t i = 0; i < n; ++i) ((double *)dst)[i] = counter_f64(seed, first + (uint64_t)i);
        return ORC_OK;
    case I32:
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t i = 0; i < n; ++i) ((int32_t *)dst)[i] = (int32_t)(uint32_t)(counter_hash(seed, first + (uint64_t)i) >> 32);
        return ORC_OK;
    case I64:
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t i = 0; i < n; ++i) ((int64_t *)dst)[i] = (int64_t)counter_hash(seed, first + (uint64_t)i);
        return ORC_OK;
    }
    return ORC_UNSUPPORTED;
}

// The parallel global sum ndarray users write, tests/par_rayon.rs:19 (and benches/par_rayon.rs):
//     a.axis_iter(Axis(0)).into_par_iter().map(|row| row.sum()).sum()
// every row's sum() is the 8-lane unrolled_fold (impl_numeric.rs:32-47 -> numeric_util.rs:14-51); rayon's `sum()` adds
// the row sums inside each task left to right and merges the tasks' subtotals in a tree whose shape is unspecified -
// restated as a static split into `threads` blocks of rows, subtotals added in block order.
template <typename T>
static int par_row_sum_typed(const orc_view *v, int threads, void *out)
{
    if (v->ndim != 2) return ORC_UNSUPPORTED;
    const int64_t rows = v->dim[0], cols = v->dim[1];
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, rows));
    std::vector<T> sub((size_t)nt, (T)0);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; ++t) {
        const int64_t lo = rows * t / nt, hi = rows * (t + 1) / nt;
        T acc = (T)0;
        for (int64_t r = lo; r < hi; ++r) {
            const T *row = (const T *)v->ptr + r * v->stride[0];
            T part;
            if (cols <= 1 || v->stride[1] == 1) {
                part = unrolled_fold<T>(row, cols, (T)0, [](T a, T b) { return wadd(a, b); });
            } else {
                part = (T)0;
                for (int64_t i = 0; i < cols; ++i) part = wadd(part, row[i * v->stride[1]]);
            }
            acc = wadd(acc, part);
        }
        sub[(size_t)t] = acc;
    }
    T total = (T)0;
    for (int t = 0; t < nt; ++t) total = wadd(total, sub[(size_t)t]);
    *(T *)out = total;
    return ORC_OK;
}

extern "C" int orc_par_row_sum(const orc_view *v, int threads, void *out)
{
    switch (v->dtype) {
    case F32: return par_row_sum_typed<float>(v, threads, out);
    case F64: return par_row_sum_typed<double>(v, threads, out);
    case I32: return par_row_sum_typed<int32_t>(v, threads, out);
    case I64: return par_row_sum_typed<int64_t>(v, threads, out);
    }
    return ORC_UNSUPPORTED;
}

// Config 5 without materialising 48 GiB: for the flat indices [first, first + n) of a = counter(seed_a), b, c this
// regenerates the inputs on the fly and returns
//   sum_f64      sum of fl32(fl32(b*c) + a) accumulated in f64 (error ~1e-16 * n: the tolerance oracle for ANY summation order)
//   abs_f64      sum of |a'|                                   (scale for the relative tolerance)
//   sum_ndarray  (only if want_ndarray_order) what ndarray's own `a.sum()` returns in f32: ONE unrolled_fold over the whole
//                block in memory order, i.e. eight interleaved sequential chains (numeric_util.rs:14-51).  The chains are
//                independent until the final combine, so each runs on its own thread - same result, 8x sooner.
extern "C" int orc_c5_sums(uint64_t seed_a, uint64_t seed_b, uint64_t seed_c, uint64_t first, int64_t n, int threads,
                           double *sum_f64, double *abs_f64, int want_ndarray_order, float *sum_ndarray)
{
    const int nt = threads > 0 ? threads : 1;
    double s = 0.0, sa = 0.0;
#pragma omp parallel for num_threads(nt) schedule(static) reduction(+ : s, sa)
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t g = first + (uint64_t)i;
        const float prod = counter_f32(seed_b, g) * counter_f32(seed_c, g);
        const float v = prod + counter_f32(seed_a, g);
        s += (double)v;
        sa += fabs((double)v);
    }
    *sum_f64 = s;
    *abs_f64 = sa;
    if (want_ndarray_order && sum_ndarray) {
        float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int64_t groups = n / 8;
#pragma omp parallel for num_threads(nt < 8 ? nt : 8) schedule(static, 1)
        for (int lane = 0; lane < 8; ++lane) {
            float acc = 0.f;
            for (int64_t k = 0; k < groups; ++k) {
                const uint64_t g = first + (uint64_t)(k * 8 + lane);
                const float prod = counter_f32(seed_b, g) * counter_f32(seed_c, g);
                acc = acc + (prod + counter_f32(seed_a, g));
            }
            p[lane] = acc;
        }
        float acc = 0.f;
        acc = acc + (p[0] + p[4]);
        acc = acc + (p[1] + p[5]);
        acc = acc + (p[2] + p[6]);
        acc = acc + (p[3] + p[7]);
        for (int64_t i = groups * 8; i < n; ++i) {
            const uint64_t g = first + (uint64_t)i;
            const float prod = counter_f32(seed_b, g) * counter_f32(seed_c, g);
            acc = acc + (prod + counter_f32(seed_a, g));
        }
        *sum_ndarray = acc;
    }
    return ORC_OK;
}

// threads this process may run on (sched affinity), whatever OMP_NUM_THREADS says: torchrun exports OMP_NUM_THREADS=1
// to every rank, which would silently turn the multi-threaded CPU arm into a single-threaded one
extern "C" int orc_hw_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

extern "C" int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
