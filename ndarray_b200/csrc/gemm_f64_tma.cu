// f64 general_mat_mul, TMA-fed variant of gemm_f64.cu's DMMA kernel (same tiling, same arithmetic, same results).
//
// Why it exists: with per-thread cp.async staging the FP64 tensor pipe tops out at 33.1 TFLOP/s; the same kernel with
// the global->shared copies removed runs at 36.0 (benchmarks/build_variant.sh, NDB_F64_DEBUG=1), i.e. the LDGSTS traffic
// through the SM's load/store path costs 8 % no matter how it is scheduled between the DMMAs.  Here one producer thread
// hands whole stages to the TMA engine (cp.async.bulk.tensor, completion on an mbarrier), so the eight DMMA warps issue
// nothing but LDS.64 + DMMA, and stages are handed over through full/empty mbarriers instead of a CTA-wide barrier.
//
// Shared-memory layout without padding: TMA cannot pad rows, and a plain [row][BK] tile would put the four rows a
// half-warp reads into the same banks.  Both operand orientations are therefore described to TMA as 3-D tensors whose
// innermost extent is 4 doubles (32 B):
//   k-contiguous operand (element (r,k) at base[r*ld + k]):  dims (4 k, rows, K/4), strides (8 B, ld*8 B, 32 B),
//       box (4, 128, BK/4)  -> smem [k/4][row][k%4]:   a half-warp's 4 rows x 4 k are 128 contiguous bytes;
//   row-contiguous operand (element (r,k) at base[k*ld + r]): dims (4 r, K, rows/4), strides (8 B, ld*8 B, 32 B),
//       box (4, BK, 128/4)  -> smem [row/4][k][row%4]: a half-warp's 4 k x 4 rows are 128 contiguous bytes.
// Every fragment LDS.64 is conflict-free, tile overhang is zero-filled by TMA's bounds check.  The split of an axis into
// (4, n/4) needs that axis to be a multiple of 4 (K for k-contiguous operands, the row count for row-contiguous ones);
// other shapes take the cp.async kernel.
//
// Reference semantics: src/linalg/impl_linalg.rs:633-643 (general_mat_mul), :512-546 (dgemm call with arbitrary strides).
#include <cuda.h>

#include <cstdlib>

#include "gemm.h"

namespace ndb {

namespace {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int CONSUMER_WARPS = 8, THREADS = 32 * (CONSUMER_WARPS + 1);
constexpr int MI = 8, NJ = 4;  // DMMA tiles per warp: 64 x 32 warp tile, warps 2 (m) x 4 (n)
constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;  // doubles
constexpr uint32_t STAGE_BYTES = (A_STAGE + B_STAGE) * sizeof(double);
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 2 * STAGES * sizeof(uint64_t);

struct TmaArgs {
    double *c;
    int64_t m, n, k, ldc;
    double alpha, beta;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// element (row, k) of a staged 128 x BK tile, k = kk + t with kk a multiple of 4
template <bool K_CONTIG>
__device__ __forceinline__ double frag(const double *st, int row, int kk, int t)
{
    if constexpr (K_CONTIG) return st[((kk >> 2) * 128 + row) * 4 + t];
    else return st[((row >> 2) * BK + kk + t) * 4 + (row & 3)];
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 1)
gemm_f64_tma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TmaArgs p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    double *tiles = reinterpret_cast<double *>(smem_raw + pad);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pad + (size_t)STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

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
    const int ktiles = (int)((p.k + BK - 1) / BK);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);                 // the producer's expect_tx arrival
            mbar_init(empty0 + 8 * s, CONSUMER_WARPS);   // one arrival per DMMA warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ---- producer: one thread feeds the TMA engine ----
        if (lane == 0) {
            for (int kt = 0; kt < ktiles; ++kt) {
                const int s = kt % STAGES;
                if (kt >= STAGES) mbar_wait(empty0 + 8 * s, (uint32_t)(kt / STAGES - 1) & 1u);
                const uint32_t a_dst = smem_u32(tiles + (size_t)s * (A_STAGE + B_STAGE));
                const uint32_t b_dst = a_dst + A_STAGE * sizeof(double);
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                if constexpr (A_KC) tma_load_3d(a_dst, &map_a, full0 + 8 * s, 0, (int)m0, kt * (BK / 4));
                else tma_load_3d(a_dst, &map_a, full0 + 8 * s, 0, kt * BK, (int)(m0 >> 2));
                if constexpr (B_KC) tma_load_3d(b_dst, &map_b, full0 + 8 * s, 0, (int)n0, kt * (BK / 4));
                else tma_load_3d(b_dst, &map_b, full0 + 8 * s, 0, kt * BK, (int)(n0 >> 2));
            }
        }
        return;
    }

    // ---- consumers: 8 warps as 2 (m) x 4 (n), 64 x 32 warp tile ----
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % STAGES;
        mbar_wait(full0 + 8 * s, (uint32_t)(kt / STAGES) & 1u);
        const double *as = tiles + (size_t)s * (A_STAGE + B_STAGE);
        const double *bs = as + A_STAGE;
        double af[2][MI], bf[2][NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) af[0][i] = frag<A_KC>(as, wm * 64 + i * 8 + g, 0, t);
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[0][j] = frag<B_KC>(bs, wn * 32 + j * 8 + g, 0, t);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            const int cur = (kk >> 2) & 1, nxt = cur ^ 1;
            if (kk + 4 < BK) {
#pragma unroll
                for (int i = 0; i < MI; ++i) af[nxt][i] = frag<A_KC>(as, wm * 64 + i * 8 + g, kk + 4, t);
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[nxt][j] = frag<B_KC>(bs, wn * 32 + j * 8 + g, kk + 4, t);
            } else {
                // the stage's last fragments are in registers: hand it back to the producer
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * s);
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
    }

    // epilogue: thread owns rows i*8+g, column pairs j*8+2t; alpha/beta fused, beta == 0 never reads C
    const bool even_ld = (p.ldc & 1) == 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t row = m0 + wm * 64 + i * 8 + g;
        if (row >= p.m) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x K operand with leading dimension ld; k_contig: element (r,k) at base[r*ld + k], else base[k*ld + r]
bool make_map(CUtensorMap *map, const double *base, int64_t rows, int64_t k, int64_t ld, bool k_contig)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (k_contig) {
        dims[0] = 4; dims[1] = (cuuint64_t)rows; dims[2] = (cuuint64_t)(k / 4);
        box[0] = 4; box[1] = 128; box[2] = BK / 4;
    } else {
        dims[0] = 4; dims[1] = (cuuint64_t)k; dims[2] = (cuuint64_t)(rows / 4);
        box[0] = 4; box[1] = BK; box[2] = 128 / 4;
    }
    strides[0] = (cuuint64_t)ld * sizeof(double);
    strides[1] = 4 * sizeof(double);
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

// Dense operands as classified by gemm_f64.cu: A is m x k (a_kc: k contiguous), B is k x n given as n "rows" of k
// (b_kc: k contiguous), C row-major with leading dimension ldc.  Returns NDB_ERR_UNSUPPORTED when the shape or the
// driver rules this path out (the caller then uses the cp.async kernel).
int gemm_f64_tma_launch(ndb_ctx *ctx, const double *a, int64_t lda, bool a_kc, const double *b, int64_t ldb, bool b_kc, double *c,
                        int64_t ldc, int64_t m, int64_t n, int64_t k, double alpha, double beta)
{
    if (getenv("NDB_GEMM_F64_CPASYNC")) return NDB_ERR_UNSUPPORTED;
    if (k < BK || m > INT32_MAX / 2 || n > INT32_MAX / 2 || k > INT32_MAX / 2) return NDB_ERR_UNSUPPORTED;
    if ((a_kc || b_kc) && (k & 3)) return NDB_ERR_UNSUPPORTED;
    if (!a_kc && (m & 3)) return NDB_ERR_UNSUPPORTED;
    if (!b_kc && (n & 3)) return NDB_ERR_UNSUPPORTED;
    if ((lda & 1) || (ldb & 1) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return NDB_ERR_UNSUPPORTED;
    CUtensorMap ma, mb;
    if (!make_map(&ma, a, m, k, lda, a_kc) || !make_map(&mb, b, n, k, ldb, b_kc)) return NDB_ERR_UNSUPPORTED;
    TmaArgs p;
    p.c = c; p.m = m; p.n = n; p.k = k; p.ldc = ldc;
    p.alpha = alpha; p.beta = beta;
    p.tiles_m = (int)((m + BM - 1) / BM);
    p.tiles_n = (int)((n + BN - 1) / BN);
    const unsigned blocks = (unsigned)(p.tiles_m * p.tiles_n);
    static bool attr_done = false;
    if (!attr_done) {
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        NDB_CUDA(cudaFuncSetAttribute(gemm_f64_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_done = true;
    }
    if (a_kc && b_kc) gemm_f64_tma<true, true><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, p);
    else if (a_kc) gemm_f64_tma<true, false><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, p);
    else if (b_kc) gemm_f64_tma<false, true><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, p);
    else gemm_f64_tma<false, false><<<blocks, THREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, p);
    NDB_CHECK_LAUNCH();
    return NDB_OK;
}

}  // namespace ndb
