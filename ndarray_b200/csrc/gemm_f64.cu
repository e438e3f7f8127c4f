// f64 general_mat_mul on the FP64 tensor cores.
//
// tcgen05 has no f64 kind (tf32 / f16 / i8 / f8f6f4 only), so double precision runs on the warp-level
// DMMA path: `mma.sync.aligned.m8n8k4.f64`, which is the native SASS shape on sm_100a (DMMA.8x8x4; ptxas
// splits m16n8k8/k16 into it).  B200's FP64 rate is 64 FMA/clk/SM on either pipe; DMMA is used because one
// instruction retires 256 FMAs from 4 registers per thread, so the register file and issue slots stop
// being the limit and a 4-stage cp.async pipeline can keep the pipe fed.
//
// This file holds the cp.async-fed kernel (the fallback of gemm_f64_tma.cu, which is ~10 % faster: see its header).
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
#ifndef NDB_F64_WARPS_M
#define NDB_F64_WARPS_M 2
#endif
#ifndef NDB_F64_BM
#define NDB_F64_BM 128
#endif
#ifndef NDB_F64_MIN_CTAS
#define NDB_F64_MIN_CTAS 1
#endif
constexpr int BM = NDB_F64_BM, BN = 128, BK = NDB_F64_BK, STAGES = NDB_F64_STAGES;
constexpr int WARPS_M = NDB_F64_WARPS_M, WARPS_N = 4, THREADS = 32 * WARPS_M * WARPS_N;
constexpr int MI = BM / (WARPS_M * 8), NJ = BN / (WARPS_N * 8);  // DMMA tiles per warp: 8 x 4 (64x32 warp tile)
#ifndef NDB_F64_DEBUG
#define NDB_F64_DEBUG 0  // experiments only (wrong results), bit mask: 1 = no global loads in the main loop, 2 = fragments loaded at kk = 0 only, 4 = no barrier
#endif
constexpr int LDK = BK + 4;   // [row][k] layout: 20 doubles per row
constexpr int LDMA = BM + 4, LDMB = BN + 4;  // [k][row] layout: 132 doubles per k
constexpr int A_STAGE = (BM * LDK > BK * LDMA) ? BM * LDK : BK * LDMA;  // doubles
constexpr int B_STAGE = (BN * LDK > BK * LDMB) ? BN * LDK : BK * LDMB;
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
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stages one operand tile with 16-byte cp.async, one chunk (2 doubles) at a time so that the main loop can slot single
// chunks between DMMA groups.  ROWS x BK logical tile of an operand whose logical element (r, k) lives at
// base[r * ld + k] (K_CONTIG) or base[k * ld + r] (!K_CONTIG); out-of-range parts are zero-filled (src-size < 16).
// Everything that does not change from one k-tile to the next is computed once per thread: chunk c = tid + i * THREADS
// sits at a fixed (row, k) offset plus i * istep, so issuing chunk i costs one 64-bit multiply-add and a select.
template <bool K_CONTIG, int ROWS>
struct TileLoader {
    static constexpr int CHUNKS = (ROWS * BK / 2) / THREADS;  // per thread per k-tile
    static_assert((ROWS * BK / 2) % THREADS == 0, "tile chunks must divide over the threads");
    static_assert(K_CONTIG ? THREADS % (BK / 2) == 0 : THREADS % (ROWS / 2) == 0, "chunk i must keep its column");
    const double *base;
    const double *src;   // chunk 0 of the k-tile to be loaded next
    int64_t istep;       // elements from chunk i to chunk i + 1
    int64_t kstep;       // elements from one k-tile to the next
    int64_t k_left;      // K_CONTIG: k elements left at this thread's k offset; else at chunk 0's k row
    int smem0;           // element offset of chunk 0 inside a stage
    uint32_t row_info;   // K_CONTIG: bit i = chunk i's row is inside the matrix; else bytes available along the row (0/8/16)

    __device__ __forceinline__ TileLoader(const double *b, int64_t ld, int64_t r0, int64_t rows_total, int64_t k_total, int tid)
        : base(b)
    {
        if constexpr (K_CONTIG) {
            const int r = tid / (BK / 2), kc = (tid % (BK / 2)) * 2;
            src = b + (r0 + r) * ld + kc;
            istep = (int64_t)(THREADS / (BK / 2)) * ld;
            kstep = BK;
            k_left = k_total - kc;
            smem0 = r * LDK + kc;
            row_info = 0;
#pragma unroll
            for (int i = 0; i < CHUNKS; ++i)
                if (r0 + r + i * (THREADS / (BK / 2)) < rows_total) row_info |= 1u << i;
        } else {
            const int kk = tid / (ROWS / 2), rc = (tid % (ROWS / 2)) * 2;
            src = b + (int64_t)kk * ld + r0 + rc;
            istep = (int64_t)(THREADS / (ROWS / 2)) * ld;
            kstep = (int64_t)BK * ld;
            k_left = k_total - kk;
            smem0 = kk * (ROWS + 4) + rc;
            const int64_t left = rows_total - (r0 + rc);
            row_info = left >= 2 ? 16u : left == 1 ? 8u : 0u;
        }
    }
    // chunk i of the current k-tile into the stage starting at `stage`
    __device__ __forceinline__ void issue(double *stage, int i) const
    {
        int bytes;
        int soff;
        if constexpr (K_CONTIG) {
            bytes = ((row_info >> i) & 1u) ? (k_left >= 2 ? 16 : k_left == 1 ? 8 : 0) : 0;
            soff = smem0 + i * (THREADS / (BK / 2)) * LDK;
        } else {
            bytes = (k_left - i * (THREADS / (ROWS / 2)) > 0) ? (int)row_info : 0;
            soff = smem0 + i * (THREADS / (ROWS / 2)) * (ROWS + 4);
        }
        cp_async16(stage + soff, bytes ? src + i * istep : base, bytes);
    }
    __device__ __forceinline__ void next_ktile()
    {
        src += kstep;
        k_left -= BK;
    }
};

template <bool K_CONTIG, int ROWS>
__device__ __forceinline__ double frag(const double *smem, int row, int k)
{
    if constexpr (K_CONTIG) return smem[row * LDK + k];
    else return smem[k * (ROWS + 4) + row];
}

// C (row-major, ldc) <- alpha * A * B + beta * C
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, NDB_F64_MIN_CTAS) gemm_f64_dmma(F64Args p)
{
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;  // WARPS_M x 4 warps
    constexpr int WTM = MI * 8, WTN = NJ * 8;

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

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int ktiles = (int)((p.k + BK - 1) / BK);
    TileLoader<A_KC, BM> la(p.a, p.lda, m0, p.m, p.k, tid);
    TileLoader<B_KC, BN> lb(p.b, p.ldb, n0, p.n, p.k, tid);
    constexpr int NA = TileLoader<A_KC, BM>::CHUNKS, NB = TileLoader<B_KC, BN>::CHUNKS;
    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) {
#pragma unroll
            for (int c = 0; c < NA; ++c) la.issue(As + s * A_STAGE, c);
#pragma unroll
            for (int c = 0; c < NB; ++c) lb.issue(Bs + s * B_STAGE, c);
            la.next_ktile();
            lb.next_ktile();
        }
        cp_async_commit();
    }
    // The two warps of an SM sub-partition leave the barrier together, so anything both of them do between DMMAs idles
    // the tensor pipe (measured: the block of 16 cp.async per thread with its address arithmetic cost 8 %).  The refill
    // of the stage freed by the barrier is therefore cut into single chunks, one after every SLOT-th group of NJ DMMAs,
    // where it disappears in the 16-cycle issue shadow of the DMMA in front of it.
    constexpr int GROUPS = (BK / 4) * MI;  // DMMA row groups per k-tile
    constexpr int SLOT = GROUPS / (NA + NB) > 0 ? GROUPS / (NA + NB) : 1;
    static_assert(SLOT * (NA + NB) <= GROUPS, "more load chunks than DMMA groups to hide them behind");
    for (int kt = 0; kt < ktiles; ++kt) {
        const int cs = kt % STAGES;
        cp_async_wait<STAGES - 2>();
        if (!(NDB_F64_DEBUG & 4)) __syncthreads();
        const double *as = As + cs * A_STAGE;
        const double *bs = Bs + cs * B_STAGE;
        const int nk = kt + STAGES - 1;
        const bool refill = !(NDB_F64_DEBUG & 1) && nk < ktiles;
        double *an = As + (nk % STAGES) * A_STAGE;
        double *bn = Bs + (nk % STAGES) * B_STAGE;
        double af[2][MI], bf[2][NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) af[0][i] = frag<A_KC, BM>(as, wm * WTM + i * 8 + g, t);
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[0][j] = frag<B_KC, BN>(bs, wn * WTN + j * 8 + g, t);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            const int cur = (kk >> 2) & 1, nxt = cur ^ 1;
            if (kk + 4 < BK && (!(NDB_F64_DEBUG & 2) || kk == 0)) {
#pragma unroll
                for (int i = 0; i < MI; ++i) af[nxt][i] = frag<A_KC, BM>(as, wm * WTM + i * 8 + g, kk + 4 + t);
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[nxt][j] = frag<B_KC, BN>(bs, wn * WTN + j * 8 + g, kk + 4 + t);
            }
#pragma unroll
            for (int i = 0; i < MI; ++i) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
                const int group = (kk / 4) * MI + i;
                if (group % SLOT == SLOT - 1 && group / SLOT < NA + NB && refill) {
                    const int c = group / SLOT;
                    if (c < NA) la.issue(an, c);
                    else lb.issue(bn, c - NA);
                }
            }
        }
        if (refill) {
            la.next_ktile();
            lb.next_ktile();
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    // epilogue: thread owns rows i*8+g, column pairs j*8+2t
    const bool even_ld = (p.ldc & 1) == 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t row = m0 + wm * WTM + i * 8 + g;
        if (row >= p.m) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int64_t col = n0 + wn * WTN + j * 8 + 2 * t;
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
    {
        const int st = gemm_f64_tma_launch(ctx, a.ptr, a.ld, a_kc, b.ptr, b.ld, b_kc, (double *)c.ptr, c.ld, g.m, g.n, g.k, alpha, beta);
        if (st != NDB_ERR_UNSUPPORTED) return st;
    }
    const unsigned blocks = (unsigned)(p.tiles_m * p.tiles_n);
    // the opt-in is per DEVICE (a process may hold one context per GPU); setting it twice from racing threads is harmless
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_dmma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_done.fetch_or(dev_bit, std::memory_order_release);
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
