// general_mat_mul / general_mat_vec_mul front-end (reference src/linalg/impl_linalg.rs:633-743) and the
// CBLAS drop-in symbols (impl_linalg.rs:466-491, :116-122).
//
// Dispatch:
//   f64, large      -> gemm_f64.cu  (FP64 tensor-core DMMA, cp.async pipeline)
//   f32, large      -> gemm_f32.cu  (3xTF32 on tcgen05 / TMEM)
//   everything else -> gemm_any below: a shared-memory tiled kernel that takes ANY signed strides and
//                      element type (ndarray's own fallback for non-BLAS types is a naive i,j,k loop,
//                      impl_linalg.rs:601-617; integers wrap).
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>
#include <atomic>

#include "gemm.h"
#include "zip_kernels.cuh"

namespace ndb {

template <typename T>
__device__ __forceinline__ T mul_add(T a, T b, T c)
{
    if constexpr (std::is_same<T, float>::value) return fmaf(a, b, c);
    else if constexpr (std::is_same<T, double>::value) return fma(a, b, c);
    else return w_add(w_mul(a, b), c);
}

// 64x64 output tile, K step 16, 256 threads, 4x4 outputs per thread.
template <typename T>
__global__ void __launch_bounds__(256) gemm_any(GemmArgs g, T alpha, T beta)
{
    pdl_enter();
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN + 1];
    const T *A = reinterpret_cast<const T *>(g.a);
    const T *B = reinterpret_cast<const T *>(g.b);
    T *Cm = reinterpret_cast<T *>(g.c);
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int tid = threadIdx.x;
    const int tr = tid / 16, tc = tid % 16;  // thread's 4x4 block: rows tr*4.., cols tc*4..
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    // loader index maps: walk whichever axis of the operand is denser with consecutive threads
    const bool a_k_fast = llabs(g.csa) <= llabs(g.rsa);
    const bool b_n_fast = llabs(g.csb) <= llabs(g.rsb);
    for (int64_t k0 = 0; k0 < g.k; k0 += BK) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int e = tid + l * 256;  // 0..1023
            int mm, kk;
            if (a_k_fast) {
                kk = e % BK;
                mm = e / BK;
            } else {
                mm = e % BM;
                kk = e / BM;
            }
            const int64_t gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < g.m && gk < g.k) ? A[gm * g.rsa + gk * g.csa] : T(0);
            int nn, kb;
            if (b_n_fast) {
                nn = e % BN;
                kb = e / BN;
            } else {
                kb = e % BK;
                nn = e / BK;
            }
            const int64_t gn = n0 + nn, gkb = k0 + kb;
            Bs[kb][nn] = (gn < g.n && gkb < g.k) ? B[gkb * g.rsb + gn * g.csb] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][tr * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tc * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = mul_add(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + tr * 4 + i;
        if (gm >= g.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gn = n0 + tc * 4 + j;
            if (gn >= g.n) continue;
            T *dst = Cm + gm * g.rsc + gn * g.csc;
            // beta == 0: C is never read (impl_linalg.rs:323-329 hands in uninitialised memory)
            if (beta == T(0)) *dst = w_mul(alpha, acc[i][j]);
            else *dst = w_add(w_mul(alpha, acc[i][j]), w_mul(beta, *dst));
        }
    }
}

// ---- matrix-vector: y <- alpha * A x + beta * y (general_mat_vec_mul, impl_linalg.rs:657-743) ------------------
// HBM-bound (A is read once).  Two forms, chosen by which axis of A is denser:
//   gemv_rows: a warp per row, lanes stride along k (coalesced when A is row-major), shuffle reduction;
//   gemv_cols_split: a thread per row, all threads walk k together (coalesced when A is column-major), k split
//   across blockIdx.y so that a few thousand rows still fill 148 SMs.
struct GemvArgs {
    const char *a, *x;
    char *y;
    int64_t m, k, rsa, csa, sx, sy;
};

template <typename T>
__global__ void __launch_bounds__(256) gemv_rows(GemvArgs g, T alpha, T beta)
{
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= g.m) return;
    const T *A = reinterpret_cast<const T *>(g.a) + row * g.rsa;
    const T *x = reinterpret_cast<const T *>(g.x);
    T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
    int64_t k = lane;
    for (; k + 96 < g.k; k += 128) {
        const T a0 = A[k * g.csa], a1 = A[(k + 32) * g.csa], a2 = A[(k + 64) * g.csa], a3 = A[(k + 96) * g.csa];
        acc0 = mul_add(a0, x[k * g.sx], acc0);
        acc1 = mul_add(a1, x[(k + 32) * g.sx], acc1);
        acc2 = mul_add(a2, x[(k + 64) * g.sx], acc2);
        acc3 = mul_add(a3, x[(k + 96) * g.sx], acc3);
    }
    for (; k < g.k; k += 32) acc0 = mul_add(A[k * g.csa], x[k * g.sx], acc0);
    T acc = w_add(w_add(acc0, acc1), w_add(acc2, acc3));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = w_add(acc, __shfl_down_sync(0xffffffffu, acc, d));
    if (lane == 0) {
        T *dst = reinterpret_cast<T *>(g.y) + row * g.sy;
        *dst = beta == T(0) ? w_mul(alpha, acc) : w_add(w_mul(alpha, acc), w_mul(beta, *dst));
    }
}

// 128-bit form of gemv_rows: A rows and x are dense and 16-byte aligned; 4 vectors of A (and of x) in flight per lane.
template <typename T>
__global__ void __launch_bounds__(256) gemv_rows_vec(GemvArgs g, T alpha, T beta)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= g.m) return;
    const V *A = reinterpret_cast<const V *>(reinterpret_cast<const T *>(g.a) + row * g.rsa);
    const V *x = reinterpret_cast<const V *>(g.x);
    const int64_t nv = g.k / N;
    T acc[4] = {T(0), T(0), T(0), T(0)};
    int64_t k = lane;
    for (; k + 96 < nv; k += 128) {
        V a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = A[k + 32 * u];
            b[u] = x[k + 32 * u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < N; ++e) acc[u] = mul_add(a[u].v[e], b[u].v[e], acc[u]);
    }
    for (; k < nv; k += 32) {
        const V a = A[k], b = x[k];
#pragma unroll
        for (int e = 0; e < N; ++e) acc[0] = mul_add(a.v[e], b.v[e], acc[0]);
    }
    T r = w_add(w_add(acc[0], acc[1]), w_add(acc[2], acc[3]));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r = w_add(r, __shfl_down_sync(0xffffffffu, r, d));
    if (lane == 0) {
        T *dst = reinterpret_cast<T *>(g.y) + row * g.sy;
        *dst = beta == T(0) ? w_mul(alpha, r) : w_add(w_mul(alpha, r), w_mul(beta, *dst));
    }
}

// Column-major A with split-K: block (x) owns 256 rows, block (y) a slice of k; partial[split][row] is merged in split
// order by gemv_cols_finish (deterministic), which also applies alpha / beta.
template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_split(GemvArgs g, int splits, T *partial)
{
    pdl_enter();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= g.m) return;
    const int64_t chunk = (g.k + splits - 1) / splits;
    const int64_t k0 = (int64_t)blockIdx.y * chunk, k1 = k0 + chunk < g.k ? k0 + chunk : g.k;
    const T *A = reinterpret_cast<const T *>(g.a) + row * g.rsa;
    const T *x = reinterpret_cast<const T *>(g.x);
    T acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = T(0);
    int64_t k = k0;
    for (; k + 7 < k1; k += 8) {
        T a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = A[(k + u) * g.csa];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = mul_add(a[u], x[(k + u) * g.sx], acc[u]);
    }
    for (; k < k1; ++k) acc[0] = mul_add(A[k * g.csa], x[k * g.sx], acc[0]);
    T r = acc[0];
#pragma unroll
    for (int u = 1; u < 8; ++u) r = w_add(r, acc[u]);
    partial[(int64_t)blockIdx.y * g.m + row] = r;
}

// 128-bit form: A is dense column-major (rsa == 1), a thread owns Vec128<T>::N consecutive rows
template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_split_vec(GemvArgs g, int splits, T *partial)
{
    pdl_enter();
    using V = Vec128<T>;
    constexpr int N = V::N;
    const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) * N;
    if (row >= g.m) return;
    const int64_t chunk = (g.k + splits - 1) / splits;
    const int64_t k0 = (int64_t)blockIdx.y * chunk, k1 = k0 + chunk < g.k ? k0 + chunk : g.k;
    const T *A = reinterpret_cast<const T *>(g.a) + row;
    const T *x = reinterpret_cast<const T *>(g.x);
    T acc[N];
#pragma unroll
    for (int e = 0; e < N; ++e) acc[e] = T(0);
    int64_t k = k0;
    for (; k + 7 < k1; k += 8) {
        V a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = *reinterpret_cast<const V *>(A + (k + u) * g.csa);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const T xv = x[(k + u) * g.sx];
#pragma unroll
            for (int e = 0; e < N; ++e) acc[e] = mul_add(a[u].v[e], xv, acc[e]);
        }
    }
    for (; k < k1; ++k) {
        const V a = *reinterpret_cast<const V *>(A + k * g.csa);
        const T xv = x[k * g.sx];
#pragma unroll
        for (int e = 0; e < N; ++e) acc[e] = mul_add(a.v[e], xv, acc[e]);
    }
#pragma unroll
    for (int e = 0; e < N; ++e) partial[(int64_t)blockIdx.y * g.m + row + e] = acc[e];
}

template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_finish(GemvArgs g, int splits, const T *partial, T alpha, T beta)
{
    pdl_enter();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= g.m) return;
    T r = partial[row];
    for (int s = 1; s < splits; ++s) r = w_add(r, partial[(int64_t)s * g.m + row]);
    T *dst = reinterpret_cast<T *>(g.y) + row * g.sy;
    *dst = beta == T(0) ? w_mul(alpha, r) : w_add(w_mul(alpha, r), w_mul(beta, *dst));
}

template <typename T>
static int gemv_launch(ndb_ctx *ctx, const GemvArgs &g, T alpha, T beta)
{
    if (g.m == 0) return NDB_OK;
    constexpr int64_t N = Vec128<T>::N;
    if (llabs(g.csa) <= llabs(g.rsa) || g.m < 2048) {
        const bool vec = g.csa == 1 && g.sx == 1 && g.k % N == 0 && g.k >= 4 * N && g.rsa % N == 0 &&
                         ((uintptr_t)g.a % 16) == 0 && ((uintptr_t)g.x % 16) == 0;
        if (vec) launch_pdl(gemv_rows_vec<T>, (unsigned)((g.m + 7) / 8), 256, 0, ctx->stream, g, alpha, beta);
        else launch_pdl(gemv_rows<T>, (unsigned)((g.m + 7) / 8), 256, 0, ctx->stream, g, alpha, beta);
        NDB_CHECK_LAUNCH();
        return NDB_OK;
    }
    const bool cvec = g.rsa == 1 && g.csa % N == 0 && g.m % N == 0 && ((uintptr_t)g.a % 16) == 0;
    const int64_t row_blocks = cvec ? (g.m / N + 255) / 256 : (g.m + 255) / 256;
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(((int64_t)ctx->sm_count * 8 + row_blocks - 1) / row_blocks,
                                                                                  std::max<int64_t>(1, g.k / 64)), 1024));
    NDB_TRY(ensure_workspace(ctx, (size_t)splits * (size_t)g.m * sizeof(T)));
    dim3 grid((unsigned)row_blocks, (unsigned)splits, 1);
    if (cvec) launch_pdl(gemv_cols_split_vec<T>, grid, 256, 0, ctx->stream, g, splits, (T *)ctx->workspace);
    else launch_pdl(gemv_cols_split<T>, grid, 256, 0, ctx->stream, g, splits, (T *)ctx->workspace);
    NDB_CHECK_LAUNCH();
    launch_pdl(gemv_cols_finish<T>, (unsigned)((g.m + 255) / 256), 256, 0, ctx->stream, g, splits, (const T *)ctx->workspace, alpha, beta);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

// C <- beta * C  (k == 0, or alpha == 0)
template <typename T>
__global__ void __launch_bounds__(256) scale_c(GemmArgs g, T beta)
{
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.m * g.n) return;
    const int64_t r = i / g.n, c = i - r * g.n;
    T *dst = reinterpret_cast<T *>(g.c) + r * g.rsc + c * g.csc;
    *dst = beta == T(0) ? T(0) : w_mul(beta, *dst);
}

template <typename T>
static int gemm_any_launch(ndb_ctx *ctx, const GemmArgs &g, T alpha, T beta)
{
    if (g.k == 0) {
        launch_pdl(scale_c<T>, (unsigned)((g.m * g.n + 255) / 256), 256, 0, ctx->stream, g, beta);
        NDB_CHECK_LAUNCH();
        return NDB_OK;
    }
    dim3 grid((unsigned)((g.n + 63) / 64), (unsigned)((g.m + 63) / 64), 1);
    if (grid.y > 65535) {
        set_error("gemm: m = %lld exceeds the fallback kernel's grid", (long long)g.m);
        return NDB_ERR_UNSUPPORTED;
    }
    launch_pdl(gemm_any<T>, grid, 256, 0, ctx->stream, g, alpha, beta);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

static int gemm_impl(ndb_ctx *ctx, const void *alpha, const ndb_view *a, const ndb_view *b, const void *beta,
                     const ndb_view *c)
{
    if (!ctx) {
        set_error("gemm: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    if (!alpha || !beta) return NDB_ERR_INVALID_ARG;
    NDB_TRY(validate_view(a, "gemm lhs"));
    NDB_TRY(validate_view(b, "gemm rhs"));
    NDB_TRY(validate_view(c, "gemm out"));
    if (a->ndim != 2 || b->ndim != 2 || c->ndim != 2) {
        set_error("general_mat_mul: operands must be 2-d");
        return NDB_ERR_INCOMPATIBLE_SHAPE;
    }
    const int64_t m = a->dim[0], k = a->dim[1], k2 = b->dim[0], n = b->dim[1];
    if (k != k2 || c->dim[0] != m || c->dim[1] != n) {
        // general_dot_shape_error (impl_linalg.rs:351-355)
        set_error("ndarray: inputs %lld × %lld, %lld × %lld, and output %lld × %lld are not compatible for matrix "
                  "multiplication",
                  (long long)m, (long long)k, (long long)k2, (long long)n, (long long)c->dim[0], (long long)c->dim[1]);
        return NDB_ERR_INCOMPATIBLE_SHAPE;
    }
    if (a->dtype != b->dtype || a->dtype != c->dtype) {
        set_error("general_mat_mul: mixed element types");
        return NDB_ERR_UNSUPPORTED;
    }
    if (m == 0 || n == 0) return NDB_OK;  // zero-sized output (tests/oper.rs:464-483)
    // the written operand may not alias itself
    if ((m > 1 && c->stride[0] == 0) || (n > 1 && c->stride[1] == 0)) {
        set_error("general_mat_mul: output has a zero stride");
        return NDB_ERR_ALIASING;
    }
    GemmArgs g;
    g.a = (const char *)a->ptr;
    g.b = (const char *)b->ptr;
    g.c = (char *)c->ptr;
    g.m = m;
    g.n = n;
    g.k = k;
    g.rsa = a->stride[0];
    g.csa = a->stride[1];
    g.rsb = b->stride[0];
    g.csb = b->stride[1];
    g.rsc = c->stride[0];
    g.csc = c->stride[1];
    DeviceGuard guard(ctx->device);
    switch (a->dtype) {
    case NDB_F64: {
        const double al = *(const double *)alpha, be = *(const double *)beta;
        if (k > 0 && gemm_f64_eligible(g)) return gemm_f64_launch(ctx, g, al, be);
        return gemm_any_launch<double>(ctx, g, al, be);
    }
    case NDB_F32: {
        const float al = *(const float *)alpha, be = *(const float *)beta;
        if (k > 0 && gemm_f32_eligible(g)) return gemm_f32_launch(ctx, g, al, be);
        return gemm_any_launch<float>(ctx, g, al, be);
    }
    case NDB_I32: return gemm_any_launch<int32_t>(ctx, g, *(const int32_t *)alpha, *(const int32_t *)beta);
    case NDB_I64: return gemm_any_launch<int64_t>(ctx, g, *(const int64_t *)alpha, *(const int64_t *)beta);
    }
    return NDB_ERR_UNSUPPORTED;
}

// ---- process-global context for the CBLAS symbols ------------------------------------------------
static std::mutex g_blas_mu;
static ndb_ctx *g_blas_ctx = nullptr;

static ndb_ctx *blas_ctx()
{
    std::lock_guard<std::mutex> lk(g_blas_mu);
    if (!g_blas_ctx) {
        int dev = 0;
        if (const char *e = getenv("NDB_DEVICE")) dev = atoi(e);
        if (ndb_ctx_create(dev, &g_blas_ctx) != NDB_OK) {
            fprintf(stderr, "libndb200: cblas entry point called but no usable B200: %s\n", ndb_last_error());
            g_blas_ctx = nullptr;
        }
    }
    return g_blas_ctx;
}

// One matrix operand as CBLAS describes it -> (rows, cols, row stride, col stride) of op(X)
struct BlasMat {
    int64_t rows, cols, rs, cs;
    int64_t span;  // elements to move
};

static BlasMat blas_operand(int layout, int trans, int64_t rows, int64_t cols, int64_t ld)
{
    // stored matrix before op(): op(X) is rows x cols
    const bool t = trans != NdbCblasNoTrans;
    const int64_t sr = t ? cols : rows, sc = t ? rows : cols;  // stored dims
    int64_t rs, cs;
    if (layout == NdbCblasRowMajor) {
        rs = ld;
        cs = 1;
    } else {
        rs = 1;
        cs = ld;
    }
    BlasMat o;
    o.rows = rows;
    o.cols = cols;
    if (t) {
        o.rs = cs;
        o.cs = rs;
    } else {
        o.rs = rs;
        o.cs = cs;
    }
    const int64_t major = layout == NdbCblasRowMajor ? sr : sc;
    const int64_t minor = layout == NdbCblasRowMajor ? sc : sr;
    o.span = major > 0 ? (major - 1) * ld + minor : 0;
    return o;
}

// Large row-major products with an untransposed A: B goes up whole, then A and C travel in row blocks while the tensor
// cores work - H2D of block i+1 (the caller's thread fills the pinned bounce buffers), the GEMM on block i (context stream)
// and D2H of block i-1 (a worker thread with its own bounce pair) overlap, so the call costs about
// upload(B) + max(GEMM, copies) instead of upload(A, B) + GEMM + download(C).
template <typename T>
static bool cblas_gemm_pipelined(ndb_ctx *ctx, int dtype, int transb, int m, int n, int k, T alpha, const T *a, int lda, const T *b, int ldb,
                                 T beta, T *c, int ldc, const BlasMat &B, bool need_c)
{
    int nb = std::max(2, std::min(8, m / 2048));
    if (getenv("NDB_CBLAS_BLOCKS")) nb = std::max(1, atoi(getenv("NDB_CBLAS_BLOCKS")));
    const int64_t rows_per = (((int64_t)m + nb - 1) / nb + 127) / 128 * 128;
    nb = (int)((m + rows_per - 1) / rows_per);
    if (prepare_bounce(ctx) != NDB_OK) return false;
    cudaStream_t s_in = ctx->copy_in, s_k = ctx->stream, s_out = ctx->copy_out;
    std::vector<cudaEvent_t> ev_in(nb), ev_k(nb);
    for (int i = 0; i < nb; ++i) {
        if (cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming) != cudaSuccess)
            return false;
    }
    cudaEvent_t ev_start, ev_b;
    cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming);
    cudaEventRecord(ev_start, s_k);  // earlier work on the context stream may still use the staging buffers
    cudaStreamWaitEvent(s_in, ev_start, 0);
    cudaStreamWaitEvent(s_out, ev_start, 0);
    bool ok = B.span == 0 || staged_h2d(ctx, ctx->stage[1], b, B.span * sizeof(T), s_in) == NDB_OK;
    cudaEventRecord(ev_b, s_in);
    cudaStreamWaitEvent(s_k, ev_b, 0);
    std::atomic<int> issued{0};
    std::atomic<bool> down_ok{true};
    const int device = ctx->device;
    std::thread drain([&] {
        cudaSetDevice(device);
        for (int i = 0; i < nb; ++i) {
            while (issued.load(std::memory_order_acquire) <= i) {
                if (issued.load(std::memory_order_acquire) < 0) return;  // the producer gave up
                std::this_thread::yield();
            }
            const int64_t r0 = (int64_t)i * rows_per, rows = std::min<int64_t>(rows_per, m - r0);
            const size_t bytes = (size_t)((rows - 1) * ldc + n) * sizeof(T);
            if (cudaStreamWaitEvent(s_out, ev_k[i], 0) != cudaSuccess ||
                staged_d2h(ctx, c + r0 * ldc, (const T *)ctx->stage[2] + r0 * ldc, bytes, s_out, 1) != NDB_OK) {
                down_ok.store(false);
                return;
            }
        }
    });
    ndb_view va, vb, vc;
    memset(&va, 0, sizeof(va));
    memset(&vb, 0, sizeof(vb));
    memset(&vc, 0, sizeof(vc));
    vb.ptr = ctx->stage[1]; vb.ndim = 2; vb.dtype = dtype; vb.dim[0] = k; vb.dim[1] = n; vb.stride[0] = B.rs; vb.stride[1] = B.cs;
    for (int i = 0; i < nb && ok; ++i) {
        const int64_t r0 = (int64_t)i * rows_per, rows = std::min<int64_t>(rows_per, m - r0);
        const size_t a_bytes = (size_t)((rows - 1) * lda + k) * sizeof(T), c_bytes = (size_t)((rows - 1) * ldc + n) * sizeof(T);
        ok = staged_h2d(ctx, (T *)ctx->stage[0] + r0 * lda, a + r0 * lda, a_bytes, s_in) == NDB_OK;
        if (ok && need_c) ok = staged_h2d(ctx, (T *)ctx->stage[2] + r0 * ldc, c + r0 * ldc, c_bytes, s_in) == NDB_OK;
        if (!ok) break;
        cudaEventRecord(ev_in[i], s_in);
        cudaStreamWaitEvent(s_k, ev_in[i], 0);
        va.ptr = (T *)ctx->stage[0] + r0 * lda; va.ndim = 2; va.dtype = dtype; va.dim[0] = rows; va.dim[1] = k; va.stride[0] = lda; va.stride[1] = 1;
        vc.ptr = (T *)ctx->stage[2] + r0 * ldc; vc.ndim = 2; vc.dtype = dtype; vc.dim[0] = rows; vc.dim[1] = n; vc.stride[0] = ldc; vc.stride[1] = 1;
        ok = gemm_impl(ctx, &alpha, &va, &vb, &beta, &vc) == NDB_OK;
        if (!ok) break;
        cudaEventRecord(ev_k[i], s_k);
        issued.store(i + 1, std::memory_order_release);
    }
    if (!ok) issued.store(-1, std::memory_order_release);
    drain.join();
    ok = ok && down_ok.load();
    if (ok) ok = cudaStreamSynchronize(s_in) == cudaSuccess && cudaStreamSynchronize(s_k) == cudaSuccess && cudaStreamSynchronize(s_out) == cudaSuccess;
    for (int i = 0; i < nb; ++i) {
        cudaEventDestroy(ev_in[i]);
        cudaEventDestroy(ev_k[i]);
    }
    cudaEventDestroy(ev_start);
    cudaEventDestroy(ev_b);
    (void)transb;
    return ok;
}

template <typename T>
static void cblas_gemm_impl(int dtype, int layout, int transa, int transb, int m, int n, int k, T alpha, const T *a, int lda,
                            const T *b, int ldb, T beta, T *c, int ldc)
{
    ndb_ctx *ctx = blas_ctx();
    if (!ctx) abort();  // a void CBLAS call has no error channel; never compute wrong results silently
    std::lock_guard<std::mutex> lk(g_blas_mu);
    DeviceGuard guard(ctx->device);
    if (m == 0 || n == 0) return;
    if (layout == NdbCblasColMajor) {
        // the same memory read row-major holds the transposes: C^T = op(B)^T op(A)^T, i.e. the row-major call with the
        // operands (and their transpose flags, and m / n) exchanged
        std::swap(a, b);
        std::swap(lda, ldb);
        std::swap(transa, transb);
        std::swap(m, n);
        layout = NdbCblasRowMajor;
    }
    const BlasMat A = blas_operand(layout, transa, m, k, lda);
    const BlasMat B = blas_operand(layout, transb, k, n, ldb);
    const BlasMat Cm = blas_operand(layout, NdbCblasNoTrans, m, n, ldc);
    bool ok = ensure_stage(ctx, 0, std::max<size_t>(16, A.span * sizeof(T))) == NDB_OK &&
              ensure_stage(ctx, 1, std::max<size_t>(16, B.span * sizeof(T))) == NDB_OK &&
              ensure_stage(ctx, 2, std::max<size_t>(16, Cm.span * sizeof(T))) == NDB_OK;
    // C travels up even for beta == 0 when it has padding (ldc > n), so the padding survives the trip back
    const bool c_dense = ldc == n;
    const bool need_c = beta != T(0) || !c_dense;
    const bool big = (size_t)A.span * sizeof(T) >= (size_t(64) << 20) && m >= 2048 && k > 0;
    if (ok && big && transa == NdbCblasNoTrans && !getenv("NDB_CBLAS_NO_PIPELINE")) {
        ok = cblas_gemm_pipelined<T>(ctx, dtype, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, B, need_c);
        if (!ok) {
            fprintf(stderr, "libndb200: cblas gemm failed: %s\n", ndb_last_error());
            abort();
        }
        return;
    }
    if (ok && A.span) ok = staged_h2d(ctx, ctx->stage[0], a, A.span * sizeof(T)) == NDB_OK;
    if (ok && B.span) ok = staged_h2d(ctx, ctx->stage[1], b, B.span * sizeof(T)) == NDB_OK;
    if (ok && need_c) ok = staged_h2d(ctx, ctx->stage[2], c, Cm.span * sizeof(T)) == NDB_OK;
    ndb_view va, vb, vc;
    memset(&va, 0, sizeof(va));
    memset(&vb, 0, sizeof(vb));
    memset(&vc, 0, sizeof(vc));
    va.ptr = ctx->stage[0]; va.ndim = 2; va.dtype = dtype; va.dim[0] = m; va.dim[1] = k; va.stride[0] = A.rs; va.stride[1] = A.cs;
    vb.ptr = ctx->stage[1]; vb.ndim = 2; vb.dtype = dtype; vb.dim[0] = k; vb.dim[1] = n; vb.stride[0] = B.rs; vb.stride[1] = B.cs;
    vc.ptr = ctx->stage[2]; vc.ndim = 2; vc.dtype = dtype; vc.dim[0] = m; vc.dim[1] = n; vc.stride[0] = Cm.rs; vc.stride[1] = Cm.cs;
    if (ok) ok = gemm_impl(ctx, &alpha, &va, &vb, &beta, &vc) == NDB_OK;
    if (ok) ok = staged_d2h(ctx, c, ctx->stage[2], Cm.span * sizeof(T)) == NDB_OK;
    if (ok) ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "libndb200: cblas gemm failed: %s\n", ndb_last_error());
        abort();
    }
}

template <typename T>
static void cblas_gemv_impl(int dtype, int layout, int trans, int m, int n, T alpha, const T *a, int lda, const T *x, int incx,
                            T beta, T *y, int incy)
{
    ndb_ctx *ctx = blas_ctx();
    if (!ctx) abort();
    std::lock_guard<std::mutex> lk(g_blas_mu);
    DeviceGuard guard(ctx->device);
    const bool t = trans != NdbCblasNoTrans;
    const int64_t ylen = t ? n : m, xlen = t ? m : n;
    if (ylen <= 0) return;
    // op(A): ylen x xlen
    const BlasMat A = blas_operand(layout, trans, ylen, xlen, lda);
    const int64_t ax = llabs((int64_t)incx), ay = llabs((int64_t)incy);
    const size_t sxb = xlen > 0 ? ((size_t)(xlen - 1) * ax + 1) * sizeof(T) : 16;
    const size_t syb = ((size_t)(ylen - 1) * ay + 1) * sizeof(T);
    bool ok = ensure_stage(ctx, 0, std::max<size_t>(16, A.span * sizeof(T))) == NDB_OK && ensure_stage(ctx, 1, sxb) == NDB_OK &&
              ensure_stage(ctx, 2, syb) == NDB_OK;
    if (ok && A.span) ok = staged_h2d(ctx, ctx->stage[0], a, A.span * sizeof(T)) == NDB_OK;
    if (ok && xlen > 0) ok = cudaMemcpyAsync(ctx->stage[1], x, sxb, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    if (ok && (beta != T(0) || ay != 1)) ok = cudaMemcpyAsync(ctx->stage[2], y, syb, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    ndb_view va, vx, vy;
    memset(&va, 0, sizeof(va));
    memset(&vx, 0, sizeof(vx));
    memset(&vy, 0, sizeof(vy));
    va.ptr = ctx->stage[0]; va.ndim = 2; va.dtype = dtype; va.dim[0] = ylen; va.dim[1] = xlen; va.stride[0] = A.rs; va.stride[1] = A.cs;
    vx.ndim = vy.ndim = 1; vx.dtype = vy.dtype = dtype;
    vx.dim[0] = xlen; vx.stride[0] = incx;
    vy.dim[0] = ylen; vy.stride[0] = incy;
    vx.ptr = (char *)ctx->stage[1] + (incx < 0 && xlen > 0 ? (int64_t)(xlen - 1) * ax * (int64_t)sizeof(T) : 0);
    vy.ptr = (char *)ctx->stage[2] + (incy < 0 ? (int64_t)(ylen - 1) * ay * (int64_t)sizeof(T) : 0);
    if (ok) ok = ndb_gemv(ctx, &alpha, &va, &vx, &beta, &vy) == NDB_OK;
    if (ok) ok = cudaMemcpyAsync(y, ctx->stage[2], syb, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
    if (ok) ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "libndb200: cblas gemv failed: %s\n", ndb_last_error());
        abort();
    }
}

template <typename T>
static T cblas_dot_impl(int dtype, int n, const T *x, int incx, const T *y, int incy)
{
    ndb_ctx *ctx = blas_ctx();
    if (!ctx) abort();
    if (n <= 0) return T(0);
    std::lock_guard<std::mutex> lk(g_blas_mu);
    DeviceGuard guard(ctx->device);
    // CBLAS passes the lowest address; a negative increment walks it from the far end (blas_1d_params,
    // impl_linalg.rs:142-154 hands ndarray's views over in exactly this form).
    const int64_t ax = llabs((int64_t)incx), ay = llabs((int64_t)incy);
    const size_t sx = ((size_t)(n - 1) * ax + 1) * sizeof(T), sy = ((size_t)(n - 1) * ay + 1) * sizeof(T);
    bool ok = ensure_stage(ctx, 0, sx) == NDB_OK && ensure_stage(ctx, 1, sy) == NDB_OK;
    if (ok) ok = cudaMemcpyAsync(ctx->stage[0], x, sx, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    if (ok) ok = cudaMemcpyAsync(ctx->stage[1], y, sy, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    ndb_view vx, vy;
    memset(&vx, 0, sizeof(vx));
    memset(&vy, 0, sizeof(vy));
    vx.ndim = vy.ndim = 1;
    vx.dtype = vy.dtype = dtype;
    vx.dim[0] = vy.dim[0] = n;
    vx.stride[0] = incx;
    vy.stride[0] = incy;
    vx.ptr = (char *)ctx->stage[0] + (incx < 0 ? (int64_t)(n - 1) * ax * (int64_t)sizeof(T) : 0);
    vy.ptr = (char *)ctx->stage[1] + (incy < 0 ? (int64_t)(n - 1) * ay * (int64_t)sizeof(T) : 0);
    T out = T(0);
    if (ok) ok = ndb_dot(ctx, &vx, &vy, &out) == NDB_OK;
    if (!ok) {
        fprintf(stderr, "libndb200: cblas dot failed: %s\n", ndb_last_error());
        abort();
    }
    return out;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_gemm(ndb_ctx *ctx, const void *alpha, const ndb_view *a, const ndb_view *b, const void *beta, const ndb_view *c)
{
    clear_error();
    return gemm_impl(ctx, alpha, a, b, beta, c);
}

// general_mat_vec_mul (impl_linalg.rs:657-743): y <- alpha*A*x + beta*y
int ndb_gemv(ndb_ctx *ctx, const void *alpha, const ndb_view *a, const ndb_view *x, const void *beta, const ndb_view *y)
{
    clear_error();
    if (!ctx) {
        set_error("gemv: null context (no CPU fallback exists)");
        return NDB_ERR_INVALID_ARG;
    }
    if (!alpha || !beta) return NDB_ERR_INVALID_ARG;
    NDB_TRY(validate_view(a, "gemv matrix"));
    NDB_TRY(validate_view(x, "gemv x"));
    NDB_TRY(validate_view(y, "gemv y"));
    if (a->ndim != 2 || x->ndim != 1 || y->ndim != 1 || a->dim[1] != x->dim[0] || a->dim[0] != y->dim[0]) {
        set_error("ndarray: inputs %lld × %lld and %lld × 1 and output %lld × 1 are not compatible for matrix multiplication",
                  (long long)(a->ndim == 2 ? a->dim[0] : -1), (long long)(a->ndim == 2 ? a->dim[1] : -1),
                  (long long)(x->ndim == 1 ? x->dim[0] : -1), (long long)(y->ndim == 1 ? y->dim[0] : -1));
        return NDB_ERR_INCOMPATIBLE_SHAPE;
    }
    if (a->dtype != x->dtype || a->dtype != y->dtype) {
        set_error("general_mat_vec_mul: mixed element types");
        return NDB_ERR_UNSUPPORTED;
    }
    if (y->dim[0] > 1 && y->stride[0] == 0) {
        set_error("general_mat_vec_mul: output has a zero stride");
        return NDB_ERR_ALIASING;
    }
    GemvArgs g;
    g.a = (const char *)a->ptr;
    g.x = (const char *)x->ptr;
    g.y = (char *)y->ptr;
    g.m = a->dim[0];
    g.k = a->dim[1];
    g.rsa = a->stride[0];
    g.csa = a->stride[1];
    g.sx = x->stride[0];
    g.sy = y->stride[0];
    DeviceGuard guard(ctx->device);
    switch (a->dtype) {
    case NDB_F32: return gemv_launch<float>(ctx, g, *(const float *)alpha, *(const float *)beta);
    case NDB_F64: return gemv_launch<double>(ctx, g, *(const double *)alpha, *(const double *)beta);
    case NDB_I32: return gemv_launch<int32_t>(ctx, g, *(const int32_t *)alpha, *(const int32_t *)beta);
    case NDB_I64: return gemv_launch<int64_t>(ctx, g, *(const int64_t *)alpha, *(const int64_t *)beta);
    }
    return NDB_ERR_UNSUPPORTED;
}

void cblas_sgemm(int layout, int transa, int transb, int m, int n, int k, float alpha, const float *a, int lda,
                 const float *b, int ldb, float beta, float *c, int ldc)
{
    cblas_gemm_impl<float>(NDB_F32, layout, transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}

void cblas_dgemm(int layout, int transa, int transb, int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc)
{
    cblas_gemm_impl<double>(NDB_F64, layout, transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}

void cblas_sgemv(int layout, int trans, int m, int n, float alpha, const float *a, int lda, const float *x, int incx,
                 float beta, float *y, int incy)
{
    cblas_gemv_impl<float>(NDB_F32, layout, trans, m, n, alpha, a, lda, x, incx, beta, y, incy);
}

void cblas_dgemv(int layout, int trans, int m, int n, double alpha, const double *a, int lda, const double *x, int incx,
                 double beta, double *y, int incy)
{
    cblas_gemv_impl<double>(NDB_F64, layout, trans, m, n, alpha, a, lda, x, incx, beta, y, incy);
}

float cblas_sdot(int n, const float *x, int incx, const float *y, int incy)
{
    return cblas_dot_impl<float>(NDB_F32, n, x, incx, y, incy);
}

double cblas_ddot(int n, const double *x, int incx, const double *y, int incy)
{
    return cblas_dot_impl<double>(NDB_F64, n, x, incx, y, incy);
}

}  // extern "C"
