// f64 general_mat_mul on the FP64 tensor cores.
//
// tcgen05 has no f64 kind (tf32 / f16 / i8 / f8f6f4 only), so double precision runs on the warp-level
// DMMA path: `mma.sync.aligned.m8n8k4.f64`, which is the native SASS shape on sm_100a (DMMA.8x8x4; ptxas
// splits m16n8k8/k16 into it).  B200's FP64 rate is 64 FMA/clk/SM on either pipe; DMMA is used because one
// instruction retires 256 FMAs from 4 registers per thread, so the register file and issue slots stop
// being the limit and a 4-stage cp.async pipeline can keep the pipe fed.
//
// Tiling: 128x128x32 block tile (3-stage ring, 216 KB; measured 32.5 vs 31.0 TFLOP/s for 128x128x16 x 4 stages),
// 256 threads = 8 warps as 2(m) x 4(n), 64x32 warp tile = 8x4 DMMA tiles,
// 64 accumulator doubles per thread.  Operands are staged with 16-byte cp.async in whichever orientation
// is dense in global memory (k-contiguous rows or m/n-contiguous columns, so C- and F-order inputs are
// both native), into shared memory padded to stride = 4 (mod 16) doubles, which makes every fragment
// LDS.64 conflict-free.  A column-major C is handled by computing C^T = B^T A^T with the same kernel.
// Epilogue fuses alpha/beta; with beta == 0, C is never read.
//
// Reference semantics: src/linalg/impl_linalg.rs:633-643 (general_mat_mul), :512-546 (dgemm call with
// arbitrary strides).  Layouts outside the fast path's rules fall back to gemm.cu's generic kernel.
#include <cstdlib>

#include "gemm.h"

extern "C" int ndb_zip(ndb_ctx *ctx, int op, const void *scalars, int n_parts, const ndb_view *parts);

namespace ndb {

namespace {

#ifndef NDB_F64_BK
#define NDB_F64_BK 32
#endif
#ifndef NDB_F64_STAGES
#define NDB_F64_STAGES 3
#endif
constexpr int BM = 128, BN = 128, BK = NDB_F64_BK, STAGES = NDB_F64_STAGES, THREADS = 256;
constexpr int LDK = BK + 4;   // [row][k] layout: 20 doubles per row
constexpr int LDM = BM + 4;   // [k][row] layout: 132 doubles per k
constexpr int A_STAGE = (BM * LDK > BK * LDM) ? BM * LDK : BK * LDM;  // doubles
constexpr int B_STAGE = (BN * LDK > BK * LDM) ? BN * LDK : BK * LDM;
constexpr size_t SMEM_BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);

struct F64Args {
    const double *a, *b;
    double *c;
    int64_t m, n, k;
    int64_t lda, ldb, ldc;  // leading (non-unit) strides, elements
    double alpha, beta;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stage one operand tile.  ROWS x BK logical tile of an operand whose logical element (r, k) lives at
// base[r * ld + k] (K_CONTIG) or base[k * ld + r] (!K_CONTIG).  Out-of-range parts are zero-filled.
template <bool K_CONTIG>
__device__ __forceinline__ void load_tile(double *smem, const double *base, int64_t ld, int64_t r0, int64_t k0,
                                          int64_t rows_total, int64_t k_total, int tid)
{
    if constexpr (K_CONTIG) {
        // 128 rows x BK/2 chunks of 2 doubles
#pragma unroll
        for (int i = 0; i < (BM * BK / 2) / THREADS; ++i) {
            const int c = tid + i * THREADS;
            const int r = c / (BK / 2), kc = (c % (BK / 2)) * 2;
            const int64_t gr = r0 + r, gk = k0 + kc;
            int bytes = 0;
            if (gr < rows_total && gk < k_total) bytes = (gk + 1 < k_total) ? 16 : 8;
            const double *src = bytes ? base + gr * ld + gk : base;
            cp_async16(smem + r * LDK + kc, src, bytes);
        }
    } else {
        // BK k x 64 chunks of 2 doubles
#pragma unroll
        for (int i = 0; i < (BM * BK / 2) / THREADS; ++i) {
            const int c = tid + i * THREADS;
            const int kk = c >> 6, rc = (c & 63) * 2;
            const int64_t gk = k0 + kk, gr = r0 + rc;
            int bytes = 0;
            if (gk < k_total && gr < rows_total) bytes = (gr + 1 < rows_total) ? 16 : 8;
            const double *src = bytes ? base + gk * ld + gr : base;
            cp_async16(smem + kk * LDM + rc, src, bytes);
        }
    }
}

template <bool K_CONTIG>
__device__ __forceinline__ double frag(const double *smem, int row, int k)
{
    if constexpr (K_CONTIG) return smem[row * LDK + k];
    else return smem[k * LDM + row];
}

// C (row-major, ldc) <- alpha * A * B + beta * C
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 1) gemm_f64_dmma(F64Args p)
{
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps

    // rasterise in groups of 8 tile-rows so that concurrently resident blocks share A and B panels in L2
    constexpr int GROUP = 8;
    const int bid = blockIdx.x;
    const int per_group = GROUP * p.tiles_n;
    const int grp = bid / per_group;
    const int first_m = grp * GROUP;
    const int gsize = min(GROUP, p.tiles_m - first_m);
    const int in_grp = bid - grp * per_group;
    const int tile_m = first_m + in_grp % gsize;
    const int tile_n = in_grp / gsize;
    const int64_t m0 = (int64_t)tile_m * BM, n0 = (int64_t)tile_n * BN;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int ktiles = (int)((p.k + BK - 1) / BK);
    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) {
            load_tile<A_KC>(As + s * A_STAGE, p.a, p.lda, m0, (int64_t)s * BK, p.m, p.k, tid);
            load_tile<B_KC>(Bs + s * B_STAGE, p.b, p.ldb, n0, (int64_t)s * BK, p.n, p.k, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < ktiles) {
                const int s = nk % STAGES;
                load_tile<A_KC>(As + s * A_STAGE, p.a, p.lda, m0, (int64_t)nk * BK, p.m, p.k, tid);
                load_tile<B_KC>(Bs + s * B_STAGE, p.b, p.ldb, n0, (int64_t)nk * BK, p.n, p.k, tid);
            }
            cp_async_commit();
        }
        const double *as = As + (kt % STAGES) * A_STAGE;
        const double *bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = frag<A_KC>(as, wm * 64 + i * 8 + g, kk + t);
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = frag<B_KC>(bs, wn * 32 + j * 8 + g, kk + t);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns rows i*8+g, column pairs j*8+2t
    const bool even_ld = (p.ldc & 1) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = m0 + wm * 64 + i * 8 + g;
        if (row >= p.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + wn * 32 + j * 8 + 2 * t;
            if (col >= p.n) continue;
            double *dst = p.c + row * p.ldc + col;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            if (col + 1 < p.n && even_ld) {
                if (p.beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2 *>(dst);
                    v0 = fma(p.beta, old.x, v0);
                    v1 = fma(p.beta, old.y, v1);
                }
                *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
            } else {
                if (p.beta != 0.0) v0 = fma(p.beta, dst[0], v0);
                dst[0] = v0;
                if (col + 1 < p.n) {
                    if (p.beta != 0.0) v1 = fma(p.beta, dst[1], v1);
                    dst[1] = v1;
                }
            }
        }
    }
}

struct Operand {
    const double *ptr;
    int64_t ld;
    bool unit_inner;  // true: element (r, c) at ptr[r*ld + c]; false: ptr[c*ld + r]
    bool ok;
};

// A dense 2-d operand with one unit stride, positive leading stride, 16-byte aligned rows.
Operand classify(const char *ptr, int64_t rows, int64_t cols, int64_t rs, int64_t cs)
{
    Operand o{(const double *)ptr, 0, true, false};
    if (((uintptr_t)ptr & 15) != 0) return o;
    if (cs == 1 && (rs >= cols || rows == 1) && rs > 0 && (rs % 2 == 0 || rows == 1)) {
        o.ld = rows == 1 ? cols + (cols & 1) : rs;
        o.unit_inner = true;
        o.ok = true;
    } else if (rs == 1 && (cs >= rows || cols == 1) && cs > 0 && (cs % 2 == 0 || cols == 1)) {
        o.ld = cols == 1 ? rows + (rows & 1) : cs;
        o.unit_inner = false;
        o.ok = true;
    }
    return o;
}

}  // namespace

bool gemm_f64_eligible(const GemmArgs &g)
{
    if (getenv("NDB_GEMM_F64_GENERIC")) return false;
    if (g.m < 64 || g.n < 64 || g.k < 16) return false;  // small problems: the generic kernel launches less
    if (g.m > INT32_MAX / 2 || g.n > INT32_MAX / 2) return false;
    return true;
}

namespace {

ndb_view dview2(const void *ptr, int64_t d0, int64_t d1, int64_t s0, int64_t s1)
{
    ndb_view v;
    memset(&v, 0, sizeof(v));
    v.ptr = const_cast<void *>(ptr);
    v.ndim = 2;
    v.dtype = NDB_F64;
    v.dim[0] = d0;
    v.dim[1] = d1;
    v.stride[0] = s0;
    v.stride[1] = s1;
    return v;
}

int launch_dense(ndb_ctx *ctx, const GemmArgs &g, double alpha, double beta)
{
    const Operand a = classify(g.a, g.m, g.k, g.rsa, g.csa);
    const Operand b = classify(g.b, g.k, g.n, g.rsb, g.csb);
    const Operand c = classify(g.c, g.m, g.n, g.rsc, g.csc);
    if (!a.ok || !b.ok || !c.ok || !c.unit_inner) {
        set_error("gemm_f64: operand layout outside the tensor-core path");
        return NDB_ERR_INCOMPATIBLE_LAYOUT;
    }
    F64Args p;
    p.a = a.ptr; p.b = b.ptr; p.c = (double *)c.ptr;
    p.m = g.m; p.n = g.n; p.k = g.k;
    p.lda = a.ld; p.ldb = b.ld; p.ldc = c.ld;
    p.alpha = alpha; p.beta = beta;
    p.tiles_m = (int)((g.m + BM - 1) / BM);
    p.tiles_n = (int)((g.n + BN - 1) / BN);
    const bool a_kc = a.unit_inner;    // A(m,k): unit inner stride = k contiguous
    const bool b_kc = !b.unit_inner;   // B(k,n): unit ROW stride = k contiguous
    const unsigned blocks = (unsigned)(p.tiles_m * p.tiles_n);
    static bool attr_done = false;
    if (!attr_done) {
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_done = true;
    }
    if (a_kc && b_kc) gemm_f64_dmma<true, true><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(p);
    else if (a_kc) gemm_f64_dmma<true, false><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(p);
    else if (b_kc) gemm_f64_dmma<false, true><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(p);
    else gemm_f64_dmma<false, false><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(p);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace

// Any signed strides are accepted (like matrixmultiply::dgemm, impl_linalg.rs:531-546): an operand that is not a
// dense, aligned C- or F-order matrix (stride-2 views, reversed axes, broadcasts, odd leading dimensions) is first
// packed into a dense row-major scratch array by the Zip engine; a C outside those rules is computed in scratch
// (pre-loaded with C when beta != 0) and scattered back.
int gemm_f64_launch(ndb_ctx *ctx, const GemmArgs &g0, double alpha, double beta)
{
    GemmArgs g = g0;
    if (llabs(g.rsc) < llabs(g.csc)) {
        // column-major C: compute C^T = B^T A^T (swap operands and their roles)
        GemmArgs t = g;
        t.a = g.b; t.rsa = g.csb; t.csa = g.rsb;
        t.b = g.a; t.rsb = g.csa; t.csb = g.rsa;
        t.rsc = g.csc; t.csc = g.rsc;
        t.m = g.n; t.n = g.m;
        g = t;
    }
    const bool a_ok = classify(g.a, g.m, g.k, g.rsa, g.csa).ok;
    const bool b_ok = classify(g.b, g.k, g.n, g.rsb, g.csb).ok;
    const Operand c_cls = classify(g.c, g.m, g.n, g.rsc, g.csc);
    const bool c_ok = c_cls.ok && c_cls.unit_inner;
    const int64_t kp = (g.k + 1) & ~(int64_t)1, np = (g.n + 1) & ~(int64_t)1;
    const size_t need = (a_ok ? 0 : (size_t)g.m * kp) + (b_ok ? 0 : (size_t)g.k * np) + (c_ok ? 0 : (size_t)g.m * np);
    double *ws = nullptr;
    if (need) NDB_CUDA(cudaMallocAsync((void **)&ws, need * sizeof(double), ctx->stream));
    double *cur = ws;
    int st = NDB_OK;
    GemmArgs d = g;
    if (!a_ok) {
        ndb_view parts[2] = {dview2(cur, g.m, g.k, kp, 1), dview2(g.a, g.m, g.k, g.rsa, g.csa)};
        st = ndb_zip(ctx, NDB_OP_COPY, nullptr, 2, parts);
        d.a = (const char *)cur; d.rsa = kp; d.csa = 1;
        cur += (size_t)g.m * kp;
    }
    if (st == NDB_OK && !b_ok) {
        ndb_view parts[2] = {dview2(cur, g.k, g.n, np, 1), dview2(g.b, g.k, g.n, g.rsb, g.csb)};
        st = ndb_zip(ctx, NDB_OP_COPY, nullptr, 2, parts);
        d.b = (const char *)cur; d.rsb = np; d.csb = 1;
        cur += (size_t)g.k * np;
    }
    if (st == NDB_OK && !c_ok) {
        if (beta != 0.0) {
            ndb_view parts[2] = {dview2(cur, g.m, g.n, np, 1), dview2(g.c, g.m, g.n, g.rsc, g.csc)};
            st = ndb_zip(ctx, NDB_OP_COPY, nullptr, 2, parts);
        }
        d.c = (char *)cur; d.rsc = np; d.csc = 1;
    }
    if (st == NDB_OK) st = launch_dense(ctx, d, alpha, beta);
    if (st == NDB_OK && !c_ok) {
        ndb_view parts[2] = {dview2(g.c, g.m, g.n, g.rsc, g.csc), dview2(d.c, g.m, g.n, np, 1)};
        st = ndb_zip(ctx, NDB_OP_COPY, nullptr, 2, parts);
    }
    if (ws) cudaFreeAsync(ws, ctx->stream);
    return st;
}

}  // namespace ndb
